// jg_torch.cpp -- PyTorch operator shim over the C-ABI: TORCH_LIBRARY(jogramop, ...).
//
// BASELINE.json's north_star asks for "a PyTorch C++/CUDA extension through a thin C-ABI layer"; SURVEY.md 8(b) names
// TORCH_LIBRARY(jogramop, ...) as the caller of the C-ABI.  This file is that extension: every operator validates its
// tensors, takes the CURRENT CUDA stream of the tensor's device, allocates the outputs with ATen and forwards to the
// corresponding jg_* entry point of libjogramop_b200.so -- no arithmetic happens here.  Engines are passed as int64
// handles (the jg_engine* of CollisionEngine.handle, or one made by jogramop::engine_from_files).
//   torch.ops.jogramop.check(engine, q[n,dof], flags, want_dist, want_reason) -> (bits int32[ceil(n/32)], dist f32[n], reason u8[n])
//   torch.ops.jogramop.check_edges(engine, qa, qb, substeps, flags, want_dist) -> (bits, dist)
//   torch.ops.jogramop.check_poses(engine, poses[n,12|7], ee_link, links int[], flags, want_dist) -> (bits, dist)
//   torch.ops.jogramop.fk(engine, q, link_ids int[]) -> f32[n, len(link_ids), 3, 4]
//   torch.ops.jogramop.engine_from_files(urdf, scene_yaml, device) -> int ; torch.ops.jogramop.engine_free(int)
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/jogramop_b200.h"

namespace {

jg_engine* eng(int64_t h) {
    TORCH_CHECK(h != 0, "jogramop: engine handle is 0");
    return reinterpret_cast<jg_engine*>(static_cast<intptr_t>(h));
}
void ok(int rc, const char* what) { TORCH_CHECK(rc == JG_OK, what, ": ", jg_last_error()); }
void want_f32_cuda(const at::Tensor& t, int64_t cols, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.dim() == 2 && t.is_contiguous() && (cols < 0 || t.size(1) == cols),
                "jogramop: ", name, " must be a contiguous float32 CUDA tensor [n, ", cols, "]");
}
void* stream_of(const at::Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

std::tuple<at::Tensor, at::Tensor, at::Tensor> op_check(int64_t engine, const at::Tensor& q, int64_t flags, bool want_dist,
                                                        bool want_reason) {
    want_f32_cuda(q, -1, "q");
    c10::cuda::CUDAGuard guard(q.device());
    const int64_t n = q.size(0);
    at::Tensor bits = at::empty({(n + 31) / 32}, q.options().dtype(at::kInt));
    at::Tensor dist = at::empty({want_dist ? n : 0}, q.options());
    at::Tensor reason = at::empty({want_reason ? n : 0}, q.options().dtype(at::kByte));
    ok(jg_check(eng(engine), q.data_ptr<float>(), n, (int)q.size(1), (uint32_t)flags, reinterpret_cast<uint32_t*>(bits.data_ptr<int>()),
                want_dist ? dist.data_ptr<float>() : nullptr, want_reason ? reason.data_ptr<uint8_t>() : nullptr, stream_of(q)),
       "jg_check");
    return {bits, dist, reason};
}

std::tuple<at::Tensor, at::Tensor> op_check_edges(int64_t engine, const at::Tensor& qa, const at::Tensor& qb, int64_t substeps,
                                                  int64_t flags, bool want_dist) {
    want_f32_cuda(qa, -1, "qa");
    want_f32_cuda(qb, qa.size(1), "qb");
    TORCH_CHECK(qa.size(0) == qb.size(0) && qa.device() == qb.device(), "jogramop: qa and qb must hold the same edges on one device");
    c10::cuda::CUDAGuard guard(qa.device());
    const int64_t m = qa.size(0);
    at::Tensor bits = at::empty({(m + 31) / 32}, qa.options().dtype(at::kInt));
    at::Tensor dist = at::empty({want_dist ? m : 0}, qa.options());
    ok(jg_check_edges(eng(engine), qa.data_ptr<float>(), qb.data_ptr<float>(), m, (int)qa.size(1), (int)substeps, (uint32_t)flags,
                      reinterpret_cast<uint32_t*>(bits.data_ptr<int>()), want_dist ? dist.data_ptr<float>() : nullptr, stream_of(qa)),
       "jg_check_edges");
    return {bits, dist};
}

std::tuple<at::Tensor, at::Tensor> op_check_poses(int64_t engine, const at::Tensor& poses, int64_t ee_link, at::IntArrayRef links,
                                                  int64_t flags, bool want_dist) {
    want_f32_cuda(poses, -1, "poses");
    TORCH_CHECK(poses.size(1) == 12 || poses.size(1) == 7, "jogramop: poses must be [n,12] (3x4 row-major) or [n,7] (pos, quat xyzw)");
    c10::cuda::CUDAGuard guard(poses.device());
    const int64_t n = poses.size(0);
    std::vector<int32_t> l(links.begin(), links.end());
    at::Tensor bits = at::empty({(n + 31) / 32}, poses.options().dtype(at::kInt));
    at::Tensor dist = at::empty({want_dist ? n : 0}, poses.options());
    ok(jg_check_poses(eng(engine), poses.data_ptr<float>(), n, poses.size(1) == 12 ? JG_POSE_MAT34 : JG_POSE_POS_QUAT_XYZW, (int)ee_link,
                      l.data(), (int)l.size(), (uint32_t)flags, reinterpret_cast<uint32_t*>(bits.data_ptr<int>()),
                      want_dist ? dist.data_ptr<float>() : nullptr, stream_of(poses)),
       "jg_check_poses");
    return {bits, dist};
}

at::Tensor op_fk(int64_t engine, const at::Tensor& q, at::IntArrayRef link_ids) {
    want_f32_cuda(q, -1, "q");
    c10::cuda::CUDAGuard guard(q.device());
    std::vector<int32_t> ids(link_ids.begin(), link_ids.end());
    at::Tensor out = at::empty({q.size(0), (int64_t)ids.size(), 3, 4}, q.options());
    ok(jg_fk(eng(engine), q.data_ptr<float>(), q.size(0), (int)q.size(1), ids.data(), (int)ids.size(), nullptr, out.data_ptr<float>(),
             stream_of(q)),
       "jg_fk");
    return out;
}

int64_t op_engine_from_files(const std::string& urdf, const std::string& scene_yaml, int64_t device) {
    jg_robot* r = jg_robot_load(urdf.c_str(), nullptr);
    TORCH_CHECK(r != nullptr, "jg_robot_load: ", jg_last_error());
    jg_scene* s = nullptr;
    if (!scene_yaml.empty()) {
        s = jg_scene_load(scene_yaml.c_str(), nullptr);
        if (!s) { jg_robot_free(r); TORCH_CHECK(false, "jg_scene_load: ", jg_last_error()); }
    }
    jg_engine* e = jg_engine_create(r, s, (int)device, nullptr);   // the engine keeps device copies: the host handles can go
    jg_robot_free(r);
    if (s) jg_scene_free(s);
    TORCH_CHECK(e != nullptr, "jg_engine_create: ", jg_last_error());
    return static_cast<int64_t>(reinterpret_cast<intptr_t>(e));
}
void op_engine_free(int64_t engine) { jg_engine_free(eng(engine)); }

}  // namespace

TORCH_LIBRARY(jogramop, m) {
    m.def("check(int engine, Tensor q, int flags, bool want_dist, bool want_reason) -> (Tensor, Tensor, Tensor)");
    m.def("check_edges(int engine, Tensor qa, Tensor qb, int substeps, int flags, bool want_dist) -> (Tensor, Tensor)");
    m.def("check_poses(int engine, Tensor poses, int ee_link, int[] links, int flags, bool want_dist) -> (Tensor, Tensor)");
    m.def("fk(int engine, Tensor q, int[] link_ids) -> Tensor");
    m.def("engine_from_files(str urdf, str scene_yaml, int device) -> int", &op_engine_from_files);
    m.def("engine_free(int engine) -> ()", &op_engine_free);
}
TORCH_LIBRARY_IMPL(jogramop, CUDA, m) {
    m.impl("check", &op_check);
    m.impl("check_edges", &op_check_edges);
    m.impl("check_poses", &op_check_poses);
    m.impl("fk", &op_fk);
}
