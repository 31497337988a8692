// Global-norm gradient clip fused into a dense Adam(+L2) update over one flat fp32 parameter buffer.
// reference: torch.nn.utils.clip_grad_norm_(params, 1) at src/organization.py:161 followed by
// torch.optim.Adam(lr, betas, weight_decay) from src/utils.py:253-254 (every row is updated every step).
// HBM roofline: 28 B per parameter per step (read w,g,m,v; write w,m,v).
#include "kernels.cuh"

namespace dmt {

__device__ __forceinline__ void sqnorm_stage1_body(const float* __restrict__ g, int64_t n,
                                                            float* __restrict__ partial, BatchRef br) {
    __shared__ float sh[32];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    float s = 0.f;
    int64_t n4 = n >> 2;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = ld4(g + 4 * i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Finishes the norm, derives the step scalars and (engine mode) reduces the batch loss and advances the step.
__device__ __forceinline__ void adam_prepare_body(const float* __restrict__ partial, int n_partial, const float* sqnorm_in,
                                    float* sqnorm_out, AdamScalars* sc, AdamHyper hp, int64_t step_by_value,
                                    int* step_dev, const float* loss_rows, const int32_t* n_targets_ptr,
                                    float* loss_out, BatchRef br) {
    __shared__ float sh[32];
    int lo, hi;
    bool active = batch_range(br, lo, hi);
    if (!active) {
        if (threadIdx.x == 0) sc->active = 0;
        return;
    }
    float s = 0.f;
    if (partial != nullptr) {
        for (int i = threadIdx.x; i < n_partial; i += blockDim.x) s += partial[i];
        s = block_sum(s, sh);
    } else if (sqnorm_in != nullptr) {
        s = sqnorm_in[0];
    }
    float l = 0.f;
    if (loss_rows != nullptr) {
        for (int i = threadIdx.x; i < hi - lo; i += blockDim.x) l += loss_rows[i];
        l = block_sum(l, sh);
    }
    if (threadIdx.x == 0) {
        int64_t t = step_by_value;
        if (step_dev != nullptr) {
            *step_dev += 1;
            t = *step_dev;
        }
        float coef = 1.f;
        if (hp.max_norm > 0.f && (partial != nullptr || sqnorm_in != nullptr)) {
            float total = sqrtf(s);
            coef = fminf(1.f, hp.max_norm / (total + 1e-6f));
        }
        double bc1 = 1.0 - pow(hp.beta1, (double)t);
        double bc2 = 1.0 - pow(hp.beta2, (double)t);
        sc->coef = coef;
        sc->step_size = (float)(hp.lr / bc1);
        sc->bc2_sqrt = (float)sqrt(bc2);
        sc->active = 1;
        if (sqnorm_out != nullptr) sqnorm_out[0] = s;
        if (loss_out != nullptr && n_targets_ptr != nullptr) loss_out[0] = l / (float)n_targets_ptr[0];
    }
}

template <bool ZERO_G>
__device__ __forceinline__ void adam_body(float* __restrict__ w, float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   const AdamScalars* __restrict__ sc, AdamHyper hp) {
    if (sc->active == 0) return;
    const float coef = sc->coef, step_size = sc->step_size, bc2_sqrt = sc->bc2_sqrt;
    const float b1 = (float)hp.beta1, b2 = (float)hp.beta2, eps = (float)hp.eps, wd = (float)hp.weight_decay;
    const float omb1 = (float)(1.0 - hp.beta1), omb2 = (float)(1.0 - hp.beta2);
    int64_t n4 = n >> 2;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 w4 = ld4(w + 4 * i), g4 = ld4(g + 4 * i), m4 = ld4(m + 4 * i), v4 = ld4(v + 4 * i);
        float* pw = &w4.x;
        float* pg = &g4.x;
        float* pm = &m4.x;
        float* pv = &v4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gg = pg[k] * coef + wd * pw[k];
            pm[k] = pm[k] + (gg - pm[k]) * omb1;  // exp_avg.lerp_(grad, 1-beta1)
            pv[k] = pv[k] * b2 + omb2 * gg * gg;
            float denom = sqrtf(pv[k]) / bc2_sqrt + eps;
            pw[k] = pw[k] - step_size * (pm[k] / denom);
        }
        st4(w + 4 * i, w4);
        st4(m + 4 * i, m4);
        st4(v + 4 * i, v4);
        if (ZERO_G) st4(g + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));  // consumed: clear for the next step
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gg = g[i] * coef + wd * w[i];
        float mm = m[i] + (gg - m[i]) * omb1;
        float vv = v[i] * b2 + omb2 * gg * gg;
        float denom = sqrtf(vv) / bc2_sqrt + eps;
        w[i] = w[i] - step_size * (mm / denom);
        m[i] = mm;
        v[i] = vv;
        if (ZERO_G) g[i] = 0.f;
    }
    (void)b1;
}

__global__ void __launch_bounds__(256) sqnorm_stage1_kernel(const float* g, int64_t n, float* partial, BatchRef br) {
    sqnorm_stage1_body(g, n, partial, br);
}
__global__ void adam_prepare_kernel(const float* partial, int n_partial, const float* sqnorm_in, float* sqnorm_out,
                                    AdamScalars* sc, AdamHyper hp, int64_t step_by_value, int* step_dev,
                                    const float* loss_rows, const int32_t* n_targets_ptr, float* loss_out,
                                    BatchRef br) {
    adam_prepare_body(partial, n_partial, sqnorm_in, sqnorm_out, sc, hp, step_by_value, step_dev, loss_rows,
                      n_targets_ptr, loss_out, br);
}
template <bool ZERO_G>
__global__ void __launch_bounds__(256) adam_kernel(float* w, float* g, float* m, float* v, int64_t n,
                                                   const AdamScalars* sc, AdamHyper hp) {
    adam_body<ZERO_G>(w, g, m, v, n, sc, hp);
}
// group forms: z = organization; every organization has its own flat buffers, scalars and step counter
__global__ void __launch_bounds__(256) sqnorm_stage1_group(const OrgDev* __restrict__ orgs, int b) {
    const OrgDev& o = orgs[blockIdx.z];
    sqnorm_stage1_body(o.G, o.n_params, o.partial, BatchRef{o.row_off, o.active, b, 0, 0});
}
__global__ void adam_prepare_group(const OrgDev* __restrict__ orgs, int b, int n_partial, AdamHyper hp) {
    const OrgDev& o = orgs[blockIdx.z];
    adam_prepare_body(o.partial, n_partial, nullptr, nullptr, o.sc, hp, 0, o.step_dev, o.loss_rows, o.t_batch_cnt + b,
                      o.loss_buf + b, BatchRef{o.row_off, o.active, b, 0, 0});
}
__global__ void __launch_bounds__(256) adam_group(const OrgDev* __restrict__ orgs, AdamHyper hp) {
    const OrgDev& o = orgs[blockIdx.z];
    adam_body<true>(o.P, o.G, o.M, o.V, o.n_params, o.sc, hp);
}

int launch_group_optim(const OrgDev* orgs, int G, int b, int64_t n_params_max, AdamHyper hp, cudaStream_t st) {
    int nb = kNormBlocks / (G > 4 ? 4 : 1);  // per-organization partial count (fixed -> deterministic)
    sqnorm_stage1_group<<<dim3(nb, 1, G), 256, 0, st>>>(orgs, b);
    DMT_LAUNCH_CHECK();
    adam_prepare_group<<<dim3(1, 1, G), 512, 0, st>>>(orgs, b, nb, hp);
    DMT_LAUNCH_CHECK();
    int64_t blocks = (n_params_max / 4 + 255) / 256;
    int64_t cap = (int64_t)kNumSMs * 16 / G;
    if (cap < 8) cap = 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_group<<<dim3((int)blocks, 1, G), 256, 0, st>>>(orgs, hp);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_sqnorm_stage1(const float* g, int64_t n, float* partial, BatchRef br, cudaStream_t st) {
    sqnorm_stage1_kernel<<<kNormBlocks, 256, 0, st>>>(g, n, partial, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_adam_prepare(const float* partial, int n_partial, const float* sqnorm_in, float* sqnorm_out,
                        AdamScalars* sc, AdamHyper hp, int64_t step_by_value, int* step_dev, const float* loss_rows,
                        const int32_t* n_targets_ptr, float* loss_out, BatchRef br, cudaStream_t st) {
    adam_prepare_kernel<<<1, 512, 0, st>>>(partial, n_partial, sqnorm_in, sqnorm_out, sc, hp, step_by_value, step_dev,
                                           loss_rows, n_targets_ptr, loss_out, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_adam(float* w, float* g, float* m, float* v, int64_t n, const AdamScalars* sc, AdamHyper hp,
                bool zero_g, cudaStream_t st) {
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    if (zero_g) adam_kernel<true><<<(int)blocks, 256, 0, st>>>(w, g, m, v, n, sc, hp);
    else adam_kernel<false><<<(int)blocks, 256, 0, st>>>(w, g, m, v, n, sc, hp);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int64_t dmt_sqnorm_scratch_floats(void) { return kNormBlocks + 16; }

int dmt_sqnorm(const float* g, int64_t n, float* out, float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0, "dmt_sqnorm: bad n");
    cudaStream_t st = as_stream(stream);
    int rc = launch_sqnorm_stage1(g, n, scratch, batch_by_value(0, 1), st);
    if (rc) return rc;
    AdamHyper hp{1e-3, 0.9, 0.999, 1e-8, 0.0, 0.f};
    AdamScalars* sc = reinterpret_cast<AdamScalars*>(scratch + kNormBlocks);
    return launch_adam_prepare(scratch, kNormBlocks, nullptr, out, sc, hp, 1, nullptr, nullptr, nullptr, nullptr,
                               batch_by_value(0, 1), st);
}

int dmt_adam_clip_step(float* w, const float* g, float* m, float* v, int64_t n, const float* sqnorm, float max_norm,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0 && step >= 1, "dmt_adam_clip_step: bad argument");
    cudaStream_t st = as_stream(stream);
    AdamHyper hp{lr, beta1, beta2, eps, weight_decay, sqnorm ? max_norm : 0.f};
    AdamScalars* sc = reinterpret_cast<AdamScalars*>(scratch);
    int rc = launch_adam_prepare(nullptr, 0, sqnorm, nullptr, sc, hp, step, nullptr, nullptr, nullptr, nullptr,
                                 batch_by_value(0, 1), st);
    if (rc) return rc;
    return launch_adam(w, const_cast<float*>(g), m, v, n, sc, hp, false, st);
}

}  // extern "C"
