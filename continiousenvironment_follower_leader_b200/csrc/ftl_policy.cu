// ftl_policy.cu -- a consumer of the simulator's observations: one fused kernel for the
// rollout's policy, a 3-layer tanh MLP (obs_dim -> 128 -> 128 -> act_dim + 1) on the fused sensorPrev matrix.
//
// Separate torch kernels (two GEMMs, a head, two tanh passes, half a dozen elementwise kernels) cost the rollout
// ~0.09 ms per step -- a quarter of the simulator's own step -- and they are serial with it by data dependence.  Here the
// observation rows are read once (float32 -> bfloat16 into shared memory), the three layers run on the tensor cores
// (mma.sync m16n8k16, bfloat16 operands, float32 accumulators) with the activations kept in shared memory, and the epilogue
// writes the action (mid + half * tanh(mu + noise * scale)) straight into the row ftl_step consumes plus the value.
//
// One persistent block per SM keeps the weights in shared memory (~100 KB) and walks over tiles of 128 envs; each of its 8
// warps owns 16 rows of the tile, so the layers of a tile need no block barrier between them.
// Build: part of libftl.so (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ftl.h"

void ftl_set_error_message(const char* msg);

namespace {

constexpr int kHid = 128;          // hidden width of both layers
constexpr int kRows = 128;         // envs per tile
constexpr int kWarps = kRows / 16;
constexpr int kHidStride = kHid + 8;   // bf16 elements per shared-memory row: (stride / 2) % 32 == 4 -> conflict-free fragments

__device__ __forceinline__ void mma_bf16(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// hidden activations: one MUFU instruction (|error| ~ 2^-11, below the bfloat16 rounding the result goes through);
// with two warps per scheduler the precise tanhf (~25 instructions, 256 per row) was three quarters of the kernel
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// rows x cols bfloat16 from global (row stride src_stride) into shared memory (row stride dst_stride), 8 elements per load
__device__ __forceinline__ void copy_rows_u4(__nv_bfloat16* dst, int dst_stride, const uint16_t* src, int src_stride, int rows,
                                             int cols, int tid, int nthreads) {
    const int per_row = cols / 8, total = rows * per_row;
    for (int e0 = tid; e0 < total; e0 += 8 * nthreads) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * nthreads;
            if (e < total) {
                const int r = e / per_row, c = (e - r * per_row) * 8;
                v[u] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * src_stride + c));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * nthreads;
            if (e < total) {
                const int r = e / per_row, c = (e - r * per_row) * 8;
                *reinterpret_cast<uint4*>(dst + (size_t)r * dst_stride + c) = v[u];
            }
        }
    }
}

// acc[nt][4] += A (16 rows of `a`, K columns) x W^T (W: [n][k] row-major = torch Linear.weight), for NT column tiles of 8
template <int NT>
__device__ __forceinline__ void layer_mma(float (*acc)[4], const __nv_bfloat16* a, int a_stride, const __nv_bfloat16* w,
                                          int w_stride, int K, int g, int t) {
    for (int k0 = 0; k0 < K; k0 += 16) {
        const uint32_t a0 = lds32(a + g * a_stride + k0 + 2 * t), a1 = lds32(a + (g + 8) * a_stride + k0 + 2 * t);
        const uint32_t a2 = lds32(a + g * a_stride + k0 + 2 * t + 8), a3 = lds32(a + (g + 8) * a_stride + k0 + 2 * t + 8);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const __nv_bfloat16* wr = w + (nt * 8 + g) * w_stride + k0 + 2 * t;
            mma_bf16(acc[nt], a0, a1, a2, a3, lds32(wr), lds32(wr + 8));
        }
    }
}

__global__ void __launch_bounds__(32 * kWarps, 1)
k_policy_mlp(const FtlMlpWeights w, const float* __restrict__ obs, const float* __restrict__ noise, int n,
             float* __restrict__ actions, float* __restrict__ values, int obs_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int D = w.obs_dim, xs = D + 8;   // (xs / 2) % 32 == 4 when D % 64 == 48 or 16...; any D % 16 == 0 keeps rows 4-byte aligned
    __nv_bfloat16* W1 = reinterpret_cast<__nv_bfloat16*>(smem);                 // [128][xs]
    __nv_bfloat16* W2 = W1 + (size_t)kHid * xs;                                 // [128][136]
    __nv_bfloat16* W3 = W2 + (size_t)kHid * kHidStride;                         // [8][136]
    __nv_bfloat16* X = W3 + (size_t)8 * kHidStride;                             // [128][xs]
    __nv_bfloat16* H = X + (size_t)kRows * xs;                                  // [128][136]
    float* B1 = reinterpret_cast<float*>(H + (size_t)kRows * kHidStride);       // [128], [128], [8]
    float* B2 = B1 + kHid;
    float* B3 = B2 + kHid;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n_out = w.act_dim + 1;
    // ---- weights -> shared memory, once per block: 16-byte loads, eight in flight per thread (a block is alone on its
    //      SM, so the copy is a chain of memory round trips unless the loads are batched) --------------------------------
    copy_rows_u4(W1, xs, w.w1, D, kHid, D, tid, blockDim.x);
    copy_rows_u4(W2, kHidStride, w.w2, kHid, kHid, kHid, tid, blockDim.x);
    for (int e = tid; e < 8 * (kHid / 2); e += blockDim.x) {
        const int r = e / (kHid / 2), c = (e - r * (kHid / 2)) * 2;
        uint32_t v = 0;
        if (r < n_out) v = *reinterpret_cast<const uint32_t*>(w.w3 + (size_t)r * kHid + c);
        *reinterpret_cast<uint32_t*>(W3 + (size_t)r * kHidStride + c) = v;
    }
    for (int e = tid; e < kHid; e += blockDim.x) { B1[e] = w.b1[e]; B2[e] = w.b2[e]; }
    if (tid < 8) B3[tid] = tid < n_out ? w.b3[tid] : 0.f;
    const int n_tiles = (n + kRows - 1) / kRows;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kRows;
        __syncthreads();   // the previous tile's readers of X are done (and the weights are in place)
        // ---- observation rows -> bfloat16 in shared memory: float4 loads, coalesced, eight in flight per thread ------
        {
            const int per_row = D / 4, total = kRows * per_row;
            for (int e0 = tid; e0 < total; e0 += 8 * blockDim.x) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int e = e0 + u * blockDim.x;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < total) {
                        const int r = e / per_row, c = (e - r * per_row) * 4;
                        if (row0 + r < n) v[u] = __ldg(reinterpret_cast<const float4*>(obs + (size_t)(row0 + r) * obs_stride + c));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int e = e0 + u * blockDim.x;
                    if (e < total) {
                        const int r = e / per_row, c = (e - r * per_row) * 4;
                        uint2 p;
                        p.x = pack_bf16(v[u].x, v[u].y);
                        p.y = pack_bf16(v[u].z, v[u].w);
                        *reinterpret_cast<uint2*>(X + (size_t)r * xs + c) = p;
                    }
                }
            }
        }
        __syncthreads();
        // ---- layer 1: [16 x D] x [D x 128], tanh, -> H (this warp's 16 rows) --------------------------------------------
        const __nv_bfloat16* xa = X + (size_t)(warp * 16) * xs;
        __nv_bfloat16* ha = H + (size_t)(warp * 16) * kHidStride;
        float acc[kHid / 8][4];
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        layer_mma<kHid / 8>(acc, xa, xs, W1, xs, D, g, t);
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) {
            const int c = nt * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(ha + g * kHidStride + c) = pack_bf16(tanh_fast(acc[nt][0] + B1[c]), tanh_fast(acc[nt][1] + B1[c + 1]));
            *reinterpret_cast<uint32_t*>(ha + (g + 8) * kHidStride + c) = pack_bf16(tanh_fast(acc[nt][2] + B1[c]), tanh_fast(acc[nt][3] + B1[c + 1]));
        }
        __syncwarp();
        // ---- layer 2: [16 x 128] x [128 x 128], tanh, in place ------------------------------------------------------------
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        layer_mma<kHid / 8>(acc, ha, kHidStride, W2, kHidStride, kHid, g, t);
        __syncwarp();   // every lane has read its fragments of H before anyone overwrites them
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) {
            const int c = nt * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(ha + g * kHidStride + c) = pack_bf16(tanh_fast(acc[nt][0] + B2[c]), tanh_fast(acc[nt][1] + B2[c + 1]));
            *reinterpret_cast<uint32_t*>(ha + (g + 8) * kHidStride + c) = pack_bf16(tanh_fast(acc[nt][2] + B2[c]), tanh_fast(acc[nt][3] + B2[c + 1]));
        }
        __syncwarp();
        // ---- head: [16 x 128] x [128 x 8]; columns < act_dim are the action mean, column act_dim the value -------------
        float out[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        layer_mma<1>(out, ha, kHidStride, W3, kHidStride, kHid, g, t);
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int row = row0 + warp * 16 + g + 8 * half;
            if (row >= n) continue;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = 2 * t + j;
                const float v = out[0][2 * half + j] + B3[c];
                if (c < w.act_dim) {
                    float mu = v;
                    if (noise) mu += noise[(size_t)row * w.act_dim + c] * w.noise_scale[c];
                    actions[(size_t)row * w.act_dim + c] = w.act_mid[c] + w.act_half[c] * tanhf(mu);
                } else if (c == w.act_dim) {
                    values[row] = v;
                }
            }
        }
    }
}

}  // namespace

extern "C" int ftl_policy_mlp(const FtlMlpWeights* w, const float* obs_dev, int32_t obs_stride, const float* noise_dev, int32_t n,
                              float* actions_dev, float* values_dev, void* cuda_stream) {
    if (!w || !obs_dev || !actions_dev || !values_dev || n < 0) { ftl_set_error_message("ftl_policy_mlp: NULL argument"); return FTL_ERR_INVALID; }
    if (w->obs_dim < 16 || w->obs_dim % 16 != 0 || w->obs_dim > 288 || w->act_dim < 1 || w->act_dim > 7 || obs_stride < w->obs_dim ||
        obs_stride % 4 != 0 || (((uintptr_t)obs_dev) & 15) != 0 || (noise_dev && !w->noise_scale)) {
        ftl_set_error_message("ftl_policy_mlp: obs_dim must be a multiple of 16 in [16, 288], act_dim in [1, 7], hidden width 128, "
                              "rows 16-byte aligned");
        return FTL_ERR_INVALID;
    }
    if (n == 0) return FTL_OK;
    const int xs = w->obs_dim + 8;
    const size_t smem = sizeof(__nv_bfloat16) * ((size_t)kHid * xs + (size_t)kHid * kHidStride + 8 * kHidStride + (size_t)kRows * xs +
                                                 (size_t)kRows * kHidStride) + sizeof(float) * (2 * kHid + 8);
    int dev = 0, sms = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e == cudaSuccess && smem > (size_t)optin) {
        ftl_set_error_message("ftl_policy_mlp: obs_dim too large for the shared memory of this device");
        return FTL_ERR_INVALID;
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_policy_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        const int tiles = (n + kRows - 1) / kRows;
        k_policy_mlp<<<tiles < sms ? tiles : sms, 32 * kWarps, smem, (cudaStream_t)cuda_stream>>>(*w, obs_dev, noise_dev, n, actions_dev,
                                                                                                 values_dev, obs_stride);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str());
        return FTL_ERR_CUDA;
    }
    return FTL_OK;
}
