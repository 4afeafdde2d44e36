// vf_torch.cpp — host-side plumbing for the hot calls: allocate the step's output tensors with the torch caching
// allocator and launch through the C-ABI of include/visfly_b200.h, all in one Python->C++ transition.
//
// This is NOT a second compute path: every function here ends in the same extern "C" entry point the ctypes binding
// (visfly_b200/_lib.py) calls; it only removes ~15 us of per-step Python overhead (six torch.empty calls and the
// ctypes marshalling of 26 arguments), which at 65 536 agents is as long as the kernel itself.  Nothing here touches
// tensor contents: allocation, pointer extraction, launch.
#include <torch/extension.h>

#include <chrono>
#include <cstdlib>

#include <c10/cuda/CUDACachingAllocator.h>
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

// A tensor over a byte range of `slab`'s storage, built directly on a TensorImpl: ~0.1 us instead of the ~0.6 us a
// dispatched view op (as_strided / narrow) costs.  The result shares the slab's storage (keeps it alive) but is an
// ordinary, non-view tensor as far as autograd is concerned — which is what a fresh kernel output should be.
// The slab itself is a bare storage from the caching allocator (current device, current stream): no tensor object and
// no dispatcher round trip for it.
inline c10::Storage alloc_slab(int64_t nbytes) {
    c10::Allocator* alloc = c10::cuda::CUDACachingAllocator::get();
    return c10::Storage(c10::Storage::use_byte_size_t(), size_t(nbytes), alloc->allocate(size_t(nbytes)), alloc, false);
}

struct Slab {
    c10::Storage storage;
    c10::DispatchKeySet keys;       // of a dense CUDA tensor on this device (taken from the step's state tensor)
};

inline at::Tensor carve(const Slab& slab, int64_t byte_offset, at::ScalarType dtype, at::IntArrayRef sizes) {
    at::Tensor t = at::detail::make_tensor<c10::TensorImpl>(c10::Storage(slab.storage), slab.keys,
                                                            c10::scalarTypeToTypeMeta(dtype));
    c10::TensorImpl* impl = t.unsafeGetTensorImpl();
    impl->set_storage_offset(byte_offset / int64_t(c10::elementSize(dtype)));
    impl->set_sizes_contiguous(sizes);
    return t;
}

// (state', obs, fifo copy | None) = one control step                          -> vf_step_fwd
// `push`: the action that arrived this step; the kernel leaves an engine-owned copy of it in the step's slab (the
// comm-delay FIFO entry, reference dynamics.py:324 `action.T.clone()`), returned as the third element.
std::tuple<at::Tensor, at::Tensor, OptTensor> step_fwd(int64_t params, int64_t substeps, int64_t integrator,
                                                       int64_t action_type, int64_t flags, const at::Tensor& state_in,
                                                       const at::Tensor& action, const OptTensor& wind,
                                                       const OptTensor& push) {
    check_f32_cuda(state_in, "state_in");
    check_f32_cuda(action, "action");
    const int64_t n = state_in.size(1);
    if (wind.has_value()) {
        check_f32_cuda(*wind, "wind");
        TORCH_CHECK(wind->numel() == 4 * n, "wind must be (n, 4)");
    }
    if (push.has_value()) {
        check_f32_cuda(*push, "fifo_push");
        TORCH_CHECK(push->numel() == 4 * n, "fifo_push must be (n, 4)");
    }
    TORCH_CHECK(state_in.dim() == 3 && state_in.size(0) == VF_STATE_PLANES && state_in.size(2) == 4,
                "state_in must be (5, n, 4)");
    TORCH_CHECK(action.numel() == 4 * n, "action must be (n, 4)");
    c10::cuda::CUDAGuard guard(state_in.device());
    auto pad = [](int64_t bytes) { return (bytes + 255) / 256 * 256; };
    const int64_t o_obs = pad(80 * n);
    const int64_t o_copy = o_obs + pad(4 * VF_OBS_FLOATS * n);
    const Slab slab{alloc_slab(push.has_value() ? o_copy + 16 * n : o_copy), state_in.key_set()};
    at::Tensor state_out = carve(slab, 0, at::kFloat, {VF_STATE_PLANES, n, 4});
    at::Tensor obs = carve(slab, o_obs, at::kFloat, {n, VF_OBS_FLOATS});
    OptTensor copy;
    if (push.has_value()) copy = carve(slab, o_copy, at::kFloat, {n, 4});
    const int rc = vf_step_fwd(reinterpret_cast<const VfParams*>(params), int(n), int(substeps), int(integrator),
                               int(action_type), unsigned(flags), state_in.data_ptr<float>(), action.data_ptr<float>(),
                               state_out.data_ptr<float>(), obs.data_ptr<float>(), nullptr,
                               static_cast<const float*>(ptr(wind)), static_cast<const float*>(ptr(push)),
                               static_cast<float*>(ptr(copy)),
                               c10::cuda::getCurrentCUDAStream(state_in.device().index()).stream());
    TORCH_CHECK(rc == 0, "visfly_b200: ", vf_last_error());
    return {state_out, obs, copy};
}

// One fused env step                                                        -> vf_env_step_fwd
// Everything that stays the same from step to step (parameter blocks, kernel variant, reset table, the device word
// of the Philox step base) is bound once; the per-step call carries the tensors that change.  At 65 536 agents the
// kernel runs ~11 us, so every microsecond of host work per step shows up in the step rate.
class EnvStepper {
  public:
    EnvStepper(int64_t params, int64_t spec, int64_t substeps, int64_t integrator, int64_t action_type, int64_t flags,
               int64_t n, OptTensor reset_table, int64_t obs_width, OptTensor step_base)
        : params_(reinterpret_cast<const VfParams*>(params)), spec_(reinterpret_cast<const VfEnvSpec*>(spec)),
          substeps_(int(substeps)), integrator_(int(integrator)), action_type_(int(action_type)),
          flags_(unsigned(flags)), n_(n), obs_width_(obs_width), reset_table_(std::move(reset_table)),
          step_base_(std::move(step_base)) {
        TORCH_CHECK(params_ && spec_ && n_ >= 0 && (obs_width_ == 13 || obs_width_ == 16), "EnvStepper: bad arguments");
        if (reset_table_.has_value()) {
            check_f32_cuda(*reset_table_, "reset_table");
            TORCH_CHECK(reset_table_->numel() == 13 * n_, "reset_table must be (n, 13)");
        }
        if (step_base_.has_value())
            TORCH_CHECK(step_base_->is_cuda() && step_base_->scalar_type() == at::kLong && step_base_->numel() == 1,
                        "step_base must be a one-element int64 CUDA tensor");
        // segments of the per-step output slab, each on a 256-byte boundary
        auto pad = [](int64_t bytes) { return (bytes + 255) / 256 * 256; };
        o_status_ = pad(80 * n_);
        o_obs_ = o_status_ + pad(16 * n_);
        o_rew_ = o_obs_ + pad(4 * obs_width_ * n_);
        o_rec_ = o_rew_ + pad(4 * n_);
        o_done_ = o_rec_ + pad(16 * n_);
        o_gate_ = o_done_ + pad(n_);
        racing_ = spec_->task == VF_TASK_RACING;
        o_copy_ = o_gate_ + (racing_ ? pad(8 * n_) : 0);     // racing: the gate index as the obs dict carries it
        sz_copy_ = pad(16 * n_);
        sz_term_ = pad(4 * obs_width_ * n_);
    }

    // returns (state', status', obs, reward, done, record, terminal obs | None, fifo copy | None, gate | None)
    std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, OptTensor, OptTensor, OptTensor>
    step(const at::Tensor& state_in, const at::Tensor& action, const at::Tensor& status_in, int64_t step_index,
         int64_t env_flags, bool want_term, int64_t host_mirror, const OptTensor& wind, const OptTensor& push,
         int64_t peer_returns) {
        check_f32_cuda(state_in, "state_in");
        check_f32_cuda(action, "action");
        TORCH_CHECK(state_in.dim() == 3 && state_in.size(0) == VF_STATE_PLANES && state_in.size(1) == n_ &&
                        state_in.size(2) == 4, "state_in must be (5, n, 4)");
        TORCH_CHECK(action.numel() == 4 * n_, "action must be (n, 4)");
        TORCH_CHECK(status_in.is_cuda() && status_in.scalar_type() == at::kInt && status_in.is_contiguous() &&
                        status_in.numel() == VF_STATUS_WORDS * n_, "status_in must be a contiguous int32 (n, 4) CUDA tensor");
        if (wind.has_value()) {
            check_f32_cuda(*wind, "wind");
            TORCH_CHECK(wind->numel() == 4 * n_, "wind must be (n, 4)");
        }
        if (push.has_value()) {
            check_f32_cuda(*push, "fifo_push");
            TORCH_CHECK(push->numel() == 4 * n_, "fifo_push must be (n, 4)");
        }
        c10::cuda::CUDAGuard guard(state_in.device());
        // one trip to the caching allocator for every output of the step
        const int64_t o_term = o_copy_ + (push.has_value() ? sz_copy_ : 0);
        const Slab slab{alloc_slab(o_term + (want_term ? sz_term_ : 0)), state_in.key_set()};
        at::Tensor state_out = carve(slab, 0, at::kFloat, {VF_STATE_PLANES, n_, 4});
        at::Tensor status = carve(slab, o_status_, at::kInt, {n_, VF_STATUS_WORDS});
        at::Tensor obs = carve(slab, o_obs_, at::kFloat, {n_, obs_width_});
        at::Tensor reward = carve(slab, o_rew_, at::kFloat, {n_});
        at::Tensor record = carve(slab, o_rec_, at::kFloat, {n_, 4});
        at::Tensor done = carve(slab, o_done_, at::kBool, {n_});
        OptTensor term, copy, gate;
        if (racing_) gate = obs_width_ == 16 ? carve(slab, o_gate_, at::kLong, {n_, 1}) : carve(slab, o_gate_, at::kLong, {n_});
        if (push.has_value()) copy = carve(slab, o_copy_, at::kFloat, {n_, 4});
        if (want_term) term = carve(slab, o_term, at::kFloat, {n_, obs_width_});
        const int rc = vf_env_step_fwd(
            params_, spec_, int(n_), substeps_, integrator_, action_type_, flags_, unsigned(env_flags),
            (unsigned long long)step_index,
            step_base_.has_value() ? reinterpret_cast<const unsigned long long*>(step_base_->data_ptr<int64_t>()) : nullptr,
            state_in.data_ptr<float>(), action.data_ptr<float>(), static_cast<const float*>(ptr(wind)),
            static_cast<const float*>(ptr(push)), static_cast<const float*>(ptr(reset_table_)),
            status_in.data_ptr<int>(), state_out.data_ptr<float>(), status.data_ptr<int>(),
            static_cast<float*>(ptr(copy)), obs.data_ptr<float>(), reward.data_ptr<float>(),
            reinterpret_cast<unsigned char*>(done.data_ptr<bool>()), record.data_ptr<float>(),
            static_cast<float*>(ptr(term)), reinterpret_cast<long long*>(ptr(gate)),
            reinterpret_cast<const VfEnvMirror*>(host_mirror), reinterpret_cast<const VfPeerScatter*>(peer_returns),
            c10::cuda::getCurrentCUDAStream(state_in.device().index()).stream());
        TORCH_CHECK(rc == 0, "visfly_b200: ", vf_last_error());
        return {state_out, status, obs, reward, done, record, term, copy, gate};
    }

    // host-cost breakdown of step() (diagnostic, tools/host_profile.py): microseconds per call of (allocation + carving),
    // (the C-ABI launch alone, outputs reused), over `iters` repetitions
    std::tuple<double, double> profile(const at::Tensor& state_in, const at::Tensor& action, const at::Tensor& status_in,
                                       int64_t iters) {
        c10::cuda::CUDAGuard guard(state_in.device());
        const auto t0 = std::chrono::steady_clock::now();
        for (int64_t i = 0; i < iters; ++i) {
            const Slab slab{alloc_slab(o_copy_ + sz_copy_ + sz_term_), state_in.key_set()};
            at::Tensor a = carve(slab, 0, at::kFloat, {VF_STATE_PLANES, n_, 4});
            at::Tensor b = carve(slab, o_status_, at::kInt, {n_, VF_STATUS_WORDS});
            at::Tensor c = carve(slab, o_obs_, at::kFloat, {n_, obs_width_});
            at::Tensor d = carve(slab, o_rew_, at::kFloat, {n_});
            at::Tensor e = carve(slab, o_rec_, at::kFloat, {n_, 4});
            at::Tensor f = carve(slab, o_done_, at::kBool, {n_});
            at::Tensor g = carve(slab, o_copy_, at::kFloat, {n_, 4});
            at::Tensor h = carve(slab, o_copy_ + sz_copy_, at::kFloat, {n_, obs_width_});
        }
        const auto t1 = std::chrono::steady_clock::now();
        const Slab slab{alloc_slab(o_copy_ + sz_copy_ + sz_term_), state_in.key_set()};
        at::Tensor so = carve(slab, 0, at::kFloat, {VF_STATE_PLANES, n_, 4});
        at::Tensor st = carve(slab, o_status_, at::kInt, {n_, VF_STATUS_WORDS});
        at::Tensor ob = carve(slab, o_obs_, at::kFloat, {n_, obs_width_});
        at::Tensor rw = carve(slab, o_rew_, at::kFloat, {n_});
        at::Tensor rc = carve(slab, o_rec_, at::kFloat, {n_, 4});
        at::Tensor dn = carve(slab, o_done_, at::kBool, {n_});
        auto stream = c10::cuda::getCurrentCUDAStream(state_in.device().index()).stream();
        for (int64_t i = 0; i < iters; ++i)
            vf_env_step_fwd(params_, spec_, int(n_), substeps_, integrator_, action_type_, flags_, 0u,
                            (unsigned long long)i, nullptr, state_in.data_ptr<float>(), action.data_ptr<float>(), nullptr,
                            nullptr, static_cast<const float*>(ptr(reset_table_)), status_in.data_ptr<int>(),
                            so.data_ptr<float>(), st.data_ptr<int>(), nullptr, ob.data_ptr<float>(), rw.data_ptr<float>(),
                            reinterpret_cast<unsigned char*>(dn.data_ptr<bool>()), rc.data_ptr<float>(), nullptr, nullptr,
                            nullptr, nullptr, stream);
        const auto t2 = std::chrono::steady_clock::now();
        const double us = 1e6 / double(iters);
        return {std::chrono::duration<double>(t1 - t0).count() * us, std::chrono::duration<double>(t2 - t1).count() * us};
    }

  private:
    const VfParams* params_;
    const VfEnvSpec* spec_;
    int substeps_, integrator_, action_type_;
    unsigned flags_;
    int64_t n_, obs_width_;
    OptTensor reset_table_, step_base_;
    bool racing_ = false;
    int64_t o_status_ = 0, o_obs_ = 0, o_rew_ = 0, o_rec_ = 0, o_done_ = 0, o_gate_ = 0, o_copy_ = 0, sz_copy_ = 0, sz_term_ = 0;
};

// Spin on the kernel's completion word (VfEnvMirror.flag) with the GIL released.
void wait_flag(int64_t flag_addr, int64_t value, int64_t timeout_us) {
    int rc;
    {
        py::gil_scoped_release release;
        rc = vf_wait_flag(reinterpret_cast<const volatile unsigned*>(flag_addr), unsigned(value), (long long)timeout_us);
    }
    TORCH_CHECK(rc == 0, "visfly_b200: ", vf_last_error());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "visfly_b200 host plumbing: output allocation + C-ABI launch in one call";
    m.def("step_fwd", &step_fwd, py::arg("params"), py::arg("substeps"), py::arg("integrator"),
          py::arg("action_type"), py::arg("flags"), py::arg("state_in"), py::arg("action"),
          py::arg("wind") = py::none(), py::arg("push") = py::none());
    py::class_<EnvStepper>(m, "EnvStepper")
        .def(py::init<int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, OptTensor, int64_t, OptTensor>())
        .def("profile", &EnvStepper::profile)
        .def("step", &EnvStepper::step, py::arg("state_in"), py::arg("action"), py::arg("status_in"),
             py::arg("step_index"), py::arg("env_flags"), py::arg("want_term"), py::arg("host_mirror"),
             py::arg("wind") = py::none(), py::arg("push") = py::none(), py::arg("peer_returns") = 0);
    m.def("wait_flag", &wait_flag, py::arg("flag_addr"), py::arg("value"), py::arg("timeout_us") = 10000000);
    m.def("abi_version", []() { return vf_abi_version(); });
}
