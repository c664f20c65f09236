// q3dr_mesh.cu -- C ABI of the image-space normals: the triangle mesh object, its
// host-side BVH build, and the batched normal-image render.  The scoring entry points that consume a mesh live in
// q3dr_api.cu next to the scoring pipeline.
#include <numeric>

#include "mesh_build.h"
#include "mesh_kernels.cuh"
#include "q3dr_internal.cuh"

namespace {

int check_mesh(const q3dr_mesh* m) {
  if (!m) return fail(Q3DR_ERR_INVALID_ARG, "mesh is NULL");
  if (!m->d_tris.p) return fail(Q3DR_ERR_INVALID_ARG, "mesh holds no device data");
  return Q3DR_OK;
}

}  // namespace

extern "C" {

void q3dr_render_opts_default(q3dr_render_opts* o) {
  if (!o) return;
  o->near_plane = 0.5f;
  o->far_plane = 1e5f;
  o->clear_argb = 0xFFFFFFFFu;
}

int q3dr_decode_normal(uint32_t argb, float out[3]) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "normal_out is NULL");
  q3dr::decode_normal_code(argb, out[0], out[1], out[2]);
  return Q3DR_OK;
}

int q3dr_mesh_create(const float* tri_vertices, const float* tri_normals, size_t nf, int device, q3dr_mesh** out) {
  if (!out) return fail(Q3DR_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!tri_vertices || nf == 0) return fail(Q3DR_ERR_INVALID_ARG, "empty mesh");
  if (nf >= (1ull << 29)) return fail(Q3DR_ERR_INVALID_ARG, "too many triangles");
  for (size_t i = 0; i < nf * 9; ++i) {
    if (!std::isfinite(tri_vertices[i])) return fail(Q3DR_ERR_INVALID_ARG, "non-finite vertex coordinate");
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return fail(Q3DR_ERR_NO_DEVICE, "no usable CUDA device: the engine has no CPU path");
  }
  if (device < 0 || device >= count) return fail(Q3DR_ERR_INVALID_ARG, "bad device index");
  Q3DR_CUDA(cudaSetDevice(device));

  q3dr::MeshHost hm;
  q3dr::build_mesh_host(tri_vertices, tri_normals, nf, &hm);
  const std::vector<float4>&nodes = hm.nodes, &tris = hm.tris, &normals = hm.normals;
  if (hm.depth + 2 > (uint32_t)q3dr::kMeshStack) return fail(Q3DR_ERR_INTERNAL, "mesh tree deeper than the traversal stack");

  std::unique_ptr<q3dr_mesh> m(new (std::nothrow) q3dr_mesh());
  if (!m) return fail(Q3DR_ERR_OUT_OF_MEMORY, "host allocation failed");
  m->device = device;
  m->n_tris = nf;
  m->n_nodes = nodes.size() / 4;
  m->depth = hm.depth;
  m->root_ref = hm.root_ref;
  m->pad = hm.pad;
  auto cleanup = [&]() {
    m->d_nodes.release();
    m->d_tris.release();
    m->d_normals.release();
  };
  int st = m->d_nodes.ensure(std::max<size_t>(nodes.size(), 4));
  if (st == Q3DR_OK) st = m->d_tris.ensure(tris.size());
  if (st == Q3DR_OK) st = m->d_normals.ensure(normals.size());
  if (st != Q3DR_OK) {
    cleanup();
    return st;
  }
  e = cudaSuccess;
  if (!nodes.empty()) e = cudaMemcpy(m->d_nodes.p, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_tris.p, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_normals.p, normals.data(), normals.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev[0]);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev[1]);
  if (e != cudaSuccess) {
    cleanup();
    return fail(Q3DR_ERR_CUDA, std::string("mesh upload: ") + cudaGetErrorString(e));
  }
  *out = m.release();
  return Q3DR_OK;
}

int q3dr_mesh_destroy(q3dr_mesh* m) {
  if (!m) return Q3DR_OK;
  cudaSetDevice(m->device);
  m->d_nodes.release();
  m->d_tris.release();
  m->d_normals.release();
  m->d_rt.release();
  m->d_image.release();
  for (auto& ev : m->ev) {
    if (ev) cudaEventDestroy(ev);
  }
  delete m;
  return Q3DR_OK;
}

int q3dr_mesh_get_info(const q3dr_mesh* m, q3dr_mesh_info* info) {
  Q3DR_TRY(check_mesh(m));
  if (!info) return fail(Q3DR_ERR_INVALID_ARG, "info is NULL");
  info->n_triangles = m->n_tris;
  info->n_nodes = m->n_nodes;
  info->tree_depth = m->depth;
  info->device = m->device;
  info->device_bytes = m->device_bytes();
  info->last_render_ms = m->last_render_ms;
  return Q3DR_OK;
}

int q3dr_mesh_render_normals(q3dr_mesh* m, const float K[4], uint32_t width, uint32_t height, const float* Rt, size_t n,
                             const q3dr_render_opts* ropts, uint32_t* out_argb) {
  Q3DR_TRY(check_mesh(m));
  if (!K || !ropts || (n && (!Rt || !out_argb))) return fail(Q3DR_ERR_INVALID_ARG, "NULL argument");
  if (width == 0 || height == 0) return fail(Q3DR_ERR_INVALID_ARG, "empty image");
  if (n == 0) return Q3DR_OK;
  Q3DR_CUDA(cudaSetDevice(m->device));
  const size_t px = (size_t)width * height;
  const size_t per_pass = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(n, 65535), (1ull << 28) / px));  // <= 1 GiB of pixels per pass
  Q3DR_TRY(m->d_rt.ensure(12 * per_pass));
  Q3DR_TRY(m->d_image.ensure(per_pass * px));
  q3dr::MeshRenderParams p;
  p.mesh.nodes = m->d_nodes.p;
  p.mesh.tris = m->d_tris.p;
  p.mesh.normals = m->d_normals.p;
  p.mesh.n_tris = (uint32_t)m->n_tris;
  p.mesh.root_ref = m->root_ref;
  p.mesh.pad = m->pad;
  p.opts.near_plane = ropts->near_plane;
  p.opts.far_plane = ropts->far_plane;
  p.opts.clear_argb = ropts->clear_argb;
  p.fx = K[0];
  p.fy = K[1];
  p.cx = K[2];
  p.cy = K[3];
  p.width = width;
  p.height = height;
  p.tiles_x = (width + 7) / 8;
  p.tiles_y = (height + 3) / 4;
  p.rt = m->d_rt.p;
  p.out = m->d_image.p;
  float ms_total = 0.f;
  for (size_t v0 = 0; v0 < n; v0 += per_pass) {
    const size_t nb = std::min(per_pass, n - v0);
    Q3DR_CUDA(cudaMemcpyAsync(m->d_rt.p, Rt + 12 * v0, nb * 12 * sizeof(float), cudaMemcpyHostToDevice, 0));
    const dim3 grid((p.tiles_x * p.tiles_y + 7) / 8, (unsigned)nb);
    Q3DR_CUDA(cudaEventRecord(m->ev[0], 0));
    q3dr::mesh_render_kernel<<<grid, 256>>>(p);
    Q3DR_CUDA(cudaGetLastError());
    Q3DR_CUDA(cudaEventRecord(m->ev[1], 0));
    Q3DR_CUDA(cudaMemcpy(out_argb + v0 * px, m->d_image.p, nb * px * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) == cudaSuccess) ms_total += ms;
  }
  m->last_render_ms = ms_total;
  return Q3DR_OK;
}

}  // extern "C"
