// device_build.h -- the GPU builder of the flattened L4QBVH (device_build.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "host_common.h"

namespace yart {

struct DeviceQbvh { // the arrays of FlatQbvh, resident in device memory (owned by the caller after a successful build)
  FlatNode* nodes;
  FlatTri* tris;
  FlatTriShade* shade;
  uint32_t n_nodes, n_tris, root, n_leaves, height, max_stack;
  double bbox_min[3], bbox_max[3];
};

// L4QBVH::new (qbvh.rs:251-361) on the device; the result is byte-identical to build_qbvh (host_qbvh.cpp).
bool build_qbvh_device(cudaStream_t stream, const yart_trimesh& mesh, DeviceQbvh& out, std::string& err);
void free_qbvh_device(DeviceQbvh& q);

} // namespace yart
