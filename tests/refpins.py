"""Digest of a displayed image (u8 RGBA/RGB) into the low-frequency form the reference-held pins use
(tools/gen_reference_pins.py): linearise the display transform, box-average to n x n blocks.

Test infrastructure only."""
import numpy as np


def linearise(u8_rgb, how):
    """Inverse of the display transform the reference applied before `(256*clamp(c,0,0.999)) as u8`
    (main.rs:710-718): "srgb" = color.rs:92-107, "gamma2" = the legacy sqrt (main.rs:684-688)."""
    c = (u8_rgb.astype(np.float64) + 0.5) / 256.0  # centre of the quantisation bin
    if how == "gamma2":
        return c * c
    if how == "srgb":
        return np.where(c <= 0.0031308 * 12.92, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    raise ValueError(how)


def block_mean(img, n):
    """n = blocks per side, or (rows, cols)."""
    ny, nx = (n, n) if isinstance(n, int) else n
    h, w = img.shape[:2]
    assert h % ny == 0 and w % nx == 0
    return img.reshape(ny, h // ny, nx, w // nx, -1).mean(axis=(1, 3))


def digest(rgba_u8, n=40, how="srgb"):
    return block_mean(linearise(rgba_u8[..., :3], how), n)


def correlation(a, b):
    a = a.reshape(-1) - a.mean()
    b = b.reshape(-1) - b.mean()
    return float((a * b).sum() / np.sqrt((a * a).sum() * (b * b).sum()))


# XYZ -> linear sRGB, the reference's E-white-adapted matrix (color.rs:209-213)
XYZ_TO_RGB = np.array([[2.6896552, -1.2758621, -0.4137931],
                       [-1.0221082, 1.9782866, 0.0438216],
                       [0.0612245, -0.2244898, 1.1632653]])


def film_digest(film, spp, n=40, how="srgb"):
    """Same digest from a film of summed XYZ samples, averaging the film over each block BEFORE the
    display clamp: unbiased at low sample counts (a noisy pixel clamps, its converged value would
    not), which is what the CPU-sized oracle renders need.  main.rs:710-712 normalisation."""
    b = block_mean(film, n) * ((720.0 - 360.0) / (106.856895 * spp))
    rgb = b @ XYZ_TO_RGB.T
    hi = float(linearise(np.array([255]), how)[0])
    return np.clip(rgb, 0.0, hi)


def load_pin(name):
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / name)
    return z["blocks"].astype(np.float64)


# cornell-box regions in 40x40 block coordinates of the 600x600 frame (15-pixel blocks)
CORNELL_REGIONS = {
    "back wall": (slice(10, 17), slice(10, 31)),
    "green wall": (slice(8, 31), slice(2, 7)),
    "red wall": (slice(8, 31), slice(33, 38)),
    "ceiling": (slice(2, 5), slice(5, 36)),
    "floor": (slice(35, 38), slice(5, 20)),
    "box front": (slice(20, 31), slice(12, 20)),
}


def region_ratios(d, ref, regions=CORNELL_REGIONS):
    """name -> (luminance ratio, per-channel ratios) of digest d over the reference digest."""
    out = {}
    for k, (r, c) in regions.items():
        a, b = d[r, c].mean(axis=(0, 1)), ref[r, c].mean(axis=(0, 1))
        out[k] = (float(a.mean() / b.mean()), a / b)
    return out


def halves(d, ref):
    """(left-half ratio, right-half ratio, block correlation) of mean-RGB luminance: the david pin's statistics
    (left = matte David + background, right = the glass instance)."""
    n = d.shape[1]
    ld, lr = d.mean(axis=-1), ref.mean(axis=-1)
    return (float(ld[:, :n // 2].mean() / lr[:, :n // 2].mean()), float(ld[:, n // 2:].mean() / lr[:, n // 2:].mean()),
            correlation(ld, lr))
