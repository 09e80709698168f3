"""Multi-GPU sharding of the render: the scene replicated on every GPU, the SAMPLE RANGE of every
pixel split across them, and the per-GPU f64 XYZ films summed onto the root with ONE collective -- the library's own
`yart_film_reduce` (an in-place ncclReduce over NVLink, csrc/device_comm.cu).  This replaces the reference's tile jobs
on a thread pool and their gather over an mpsc channel (main.rs:629-646, 746-760).

Sample sharding is exact by construction: sanitize_sample_xyz acts per sample (reference main.rs:700-707), so the film
is a plain sum over samples and Philox counters make every sample independent of who renders it.  The closest-hit
sweep shards its ray array instead and needs no collective at all.

Two shapes of host:
  * one process per GPU (bench.py under torchrun): `shard_range` / `step_sample_range` + `Comm.from_id`;
  * one process driving N GPUs (yart_cli.py --gpus N): `MultiGpuRenderer` below -- N contexts, one host thread per
    context while rendering (a yart_ctx is not thread-safe, but distinct contexts are independent), `Comm.from_contexts`.
`reduce_film` / `render_distributed` are the torch.distributed flavour used by the CPU (gloo) test of the partition.
"""
import threading


def shard_range(begin, end, rank, world):
    """Contiguous, balanced split of [begin, end) -- rank r gets the r-th piece (sizes differ by <= 1)."""
    n = max(0, end - begin)
    base, extra = divmod(n, world)
    lo = begin + rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def step_sample_range(step, rank, world, spp_per_step):
    """Weak-scaling schedule of bench.py: in every step each rank renders its own spp_per_step samples."""
    lo = (step * world + rank) * spp_per_step
    return lo, lo + spp_per_step


def reduce_film(film, dst=0, group=None):
    """Sum the per-rank films onto rank `dst` (in place) with torch.distributed.  `film` is a torch tensor on the
    device the process group was created for (CPU tensors with gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def render_distributed(render_fn, film, sample_begin, sample_end, rank, world, dst=0, group=None):
    """Render this rank's share of [sample_begin, sample_end) into `film` (a zeroed torch tensor of shape
    (H, W, 3), f64) with render_fn(lo, hi, film) and combine on rank dst.  Returns this rank's (lo, hi)."""
    lo, hi = shard_range(sample_begin, sample_end, rank, world)
    if hi > lo:
        render_fn(lo, hi, film)
    reduce_film(film, dst=dst, group=group)
    return lo, hi


class MultiGpuRenderer:
    """N GPUs of one box behind the interface of a single Context: set a scene once, then `render` sample ranges (each
    GPU takes its shard of every range) and read the combined film on demand.

    Progressive use is supported: renders accumulate into per-GPU device films; `film()` reduces them onto GPU 0 in
    place and then clears the other GPUs' films, so the root keeps the running total and the next reduce adds only
    what was rendered since (sums are associative; the film differs from the 1-GPU film only in the order of f64
    additions, ~1e-16 relative)."""

    def __init__(self, pkg, devices):
        self.pkg = pkg
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("no devices")
        self.contexts = [pkg.Context(d) for d in self.devices]
        self.comm = pkg.Comm.from_contexts(self.contexts) if len(self.contexts) > 1 else None
        self.films = None
        self.shape = None
        self._base = None

    @property
    def n(self):
        return len(self.contexts)

    def _each(self, fn):
        """fn(rank, ctx) on every context, one host thread per context; re-raises the first failure."""
        errs, outs = [], [None] * self.n

        def run(r):
            try:
                outs[r] = fn(r, self.contexts[r])
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        if self.n == 1:
            run(0)
        else:
            threads = [threading.Thread(target=run, args=(r,)) for r in range(self.n)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errs:
            raise errs[0]
        return outs

    def set_scene(self, scene):
        self._each(lambda r, c: c.set_scene(scene))

    def _ensure_films(self, width, height):
        if self.shape != (width, height):
            self.release_films()
            self.films = [c.film_create(width, height) for c in self.contexts]
            self.shape = (width, height)

    def load_film(self, film):
        """Start from an existing film (resuming a checkpoint): it is kept on the host and added when the combined
        film is read; the device films start from zero."""
        import numpy as np
        h, w = film.shape[:2]
        self._ensure_films(w, h)
        for r, c in enumerate(self.contexts):
            c.film_clear(self.films[r], w, h)
        self._base = np.array(film, dtype=np.float64, copy=True)

    def render(self, camera, width, height, sample_begin, sample_end, max_depth=50, seed=1, order=None, flags=0,
               batch_spp=0):
        """Every GPU renders its shard of samples [sample_begin, sample_end) of every pixel.  Returns the per-GPU stats."""
        order = self.pkg.ORDER_NEAR if order is None else order
        self._ensure_films(width, height)

        def work(r, c):
            lo, hi = shard_range(sample_begin, sample_end, r, self.n)
            if hi <= lo:
                return None
            return c.render_device(camera, width, height, lo, hi, self.films[r], max_depth, seed, order, batch_spp, flags=flags)

        return [s for s in self._each(work) if s is not None]

    def _reduce(self):
        w, h = self.shape
        if self.comm is not None:
            self.comm.film_reduce(self.films, w, h, root=0)
            for r in range(1, self.n):  # the root now holds the total: the others start again from zero
                self.contexts[r].film_clear(self.films[r], w, h)

    def film(self):
        """The combined (H, W, 3) f64 film on the host."""
        w, h = self.shape
        self._reduce()
        out = self.contexts[0].film_read(self.films[0], w, h)
        if self._base is not None:
            out += self._base
        return out

    def finalize(self, spp):
        """RGBA8 image of the combined film (reference main.rs:710-718)."""
        return self.contexts[0].film_finalize(self.film(), spp)

    def release_films(self):
        if self.films:
            for c, f in zip(self.contexts, self.films):
                c.film_destroy(f)
        self.films, self.shape = None, None

    def close(self):
        self.release_films()
        if self.comm is not None:
            self.comm.close()
            self.comm = None
        for c in self.contexts:
            c.close()
        self.contexts = []
