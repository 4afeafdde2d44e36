// vf_torch.cpp — host-side plumbing for the hot calls: allocate the step's output tensors with the torch caching
// allocator and launch through the C-ABI of include/visfly_b200.h, all in one Python->C++ transition.
//
// This is NOT a second compute path: every function here ends in the same extern "C" entry point the ctypes binding
// (visfly_b200/_lib.py) calls; it only removes ~15 us of per-step Python overhead (six torch.empty calls and the
// ctypes marshalling of 26 arguments), which at 65 536 agents is as long as the kernel itself.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "../../include/visfly_b200.h"

namespace {

using OptTensor = c10::optional<at::Tensor>;

inline void* ptr(const OptTensor& t) { return t.has_value() ? t->data_ptr() : nullptr; }

inline void check_f32_cuda(const at::Tensor& t, const char* what) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), what,
                " must be a contiguous float32 CUDA tensor");
}

// (state', obs) = one control step                                           -> vf_step_fwd
std::tuple<at::Tensor, at::Tensor> step_fwd(int64_t params, int64_t substeps, int64_t integrator, int64_t action_type,
                                            int64_t flags, const at::Tensor& state_in, const at::Tensor& action) {
    check_f32_cuda(state_in, "state_in");
    check_f32_cuda(action, "action");
    const int64_t n = state_in.size(1);
    TORCH_CHECK(state_in.dim() == 3 && state_in.size(0) == VF_STATE_PLANES && state_in.size(2) == 4,
                "state_in must be (5, n, 4)");
    TORCH_CHECK(action.numel() == 4 * n, "action must be (n, 4)");
    c10::cuda::CUDAGuard guard(state_in.device());
    at::Tensor state_out = at::empty_like(state_in);
    at::Tensor obs = at::empty({n, VF_OBS_FLOATS}, state_in.options());
    const int rc = vf_step_fwd(reinterpret_cast<const VfParams*>(params), int(n), int(substeps), int(integrator),
                               int(action_type), unsigned(flags), state_in.data_ptr<float>(), action.data_ptr<float>(),
                               state_out.data_ptr<float>(), obs.data_ptr<float>(), nullptr,
                               c10::cuda::getCurrentCUDAStream(state_in.device().index()).stream());
    TORCH_CHECK(rc == 0, "visfly_b200: ", vf_last_error());
    return {state_out, obs};
}

// one fused env step                                                        -> vf_env_step_fwd
// returns (state', obs, reward, done, record, terminal obs | None, saved | None)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, OptTensor, OptTensor>
env_step_fwd(int64_t params, int64_t spec, int64_t substeps, int64_t integrator, int64_t action_type, int64_t flags,
             int64_t env_flags, int64_t step_index, const at::Tensor& state_in, const at::Tensor& action,
             const OptTensor& reset_table, at::Tensor& step_count, at::Tensor& returns, at::Tensor& ebits,
             const OptTensor& gate, const OptTensor& gates_passed, int64_t obs_width, bool want_term, bool want_saved,
             int64_t host_mirror) {
    check_f32_cuda(state_in, "state_in");
    check_f32_cuda(action, "action");
    const int64_t n = state_in.size(1);
    TORCH_CHECK(state_in.dim() == 3 && state_in.size(0) == VF_STATE_PLANES && state_in.size(2) == 4,
                "state_in must be (5, n, 4)");
    TORCH_CHECK(action.numel() == 4 * n, "action must be (n, 4)");
    if (reset_table.has_value()) check_f32_cuda(*reset_table, "reset_table");
    check_f32_cuda(returns, "returns");
    TORCH_CHECK(step_count.is_cuda() && step_count.scalar_type() == at::kInt && step_count.is_contiguous() &&
                    step_count.numel() == n, "step_count must be a contiguous int32 CUDA tensor of n elements");
    TORCH_CHECK(ebits.is_cuda() && ebits.scalar_type() == at::kByte && ebits.is_contiguous() && ebits.numel() == n,
                "ebits must be a contiguous uint8 CUDA tensor of n elements");
    for (const OptTensor* t : {&gate, &gates_passed})
        TORCH_CHECK(!t->has_value() || ((*t)->is_cuda() && (*t)->scalar_type() == at::kInt && (*t)->is_contiguous() &&
                                        (*t)->numel() == n), "gate / gates_passed must be contiguous int32 CUDA tensors");
    c10::cuda::CUDAGuard guard(state_in.device());
    const auto f32 = state_in.options();
    // one trip to the caching allocator for all float outputs (each allocation costs ~1 us of host time, the kernel
    // 13 us): segments of one slab, each starting on a 256-byte boundary; the views keep the slab alive
    auto pad = [](int64_t floats) { return (floats + 63) / 64 * 64; };
    const int64_t o_state = 0, o_obs = o_state + pad(20 * n), o_rew = o_obs + pad(obs_width * n),
                  o_rec = o_rew + pad(n), o_term = o_rec + pad(4 * n),
                  total = o_term + (want_term ? pad(obs_width * n) : 0);
    at::Tensor slab = at::empty({total}, f32);
    at::Tensor state_out = slab.as_strided({VF_STATE_PLANES, n, 4}, {4 * n, 4, 1}, o_state);
    at::Tensor obs = slab.as_strided({n, obs_width}, {obs_width, 1}, o_obs);
    at::Tensor reward = slab.as_strided({n}, {1}, o_rew);
    at::Tensor record = slab.as_strided({n, 4}, {4, 1}, o_rec);
    at::Tensor done = at::empty({n}, f32.dtype(at::kBool));
    OptTensor term, saved;
    if (want_term) term = slab.as_strided({n, obs_width}, {obs_width, 1}, o_term);
    if (want_saved) saved = at::empty({n, 2}, f32.dtype(at::kInt));
    const int rc = vf_env_step_fwd(
        reinterpret_cast<const VfParams*>(params), reinterpret_cast<const VfEnvSpec*>(spec), int(n), int(substeps),
        int(integrator), int(action_type), unsigned(flags), unsigned(env_flags), (unsigned long long)step_index,
        state_in.data_ptr<float>(), action.data_ptr<float>(), static_cast<const float*>(ptr(reset_table)),
        step_count.data_ptr<int>(), returns.data_ptr<float>(), ebits.data_ptr<uint8_t>(),
        static_cast<int*>(ptr(gate)), static_cast<int*>(ptr(gates_passed)), state_out.data_ptr<float>(),
        obs.data_ptr<float>(), reward.data_ptr<float>(), reinterpret_cast<unsigned char*>(done.data_ptr<bool>()),
        record.data_ptr<float>(), static_cast<float*>(ptr(term)), static_cast<int*>(ptr(saved)),
        reinterpret_cast<const VfEnvMirror*>(host_mirror),
        c10::cuda::getCurrentCUDAStream(state_in.device().index()).stream());
    TORCH_CHECK(rc == 0, "visfly_b200: ", vf_last_error());
    return {state_out, obs, reward, done, record, term, saved};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "visfly_b200 host plumbing: output allocation + C-ABI launch in one call";
    m.def("step_fwd", &step_fwd);
    m.def("env_step_fwd", &env_step_fwd);
    m.def("abi_version", []() { return vf_abi_version(); });
}
