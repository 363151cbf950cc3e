// ftl_policy.cu -- a consumer of the simulator's observations: one fused kernel for the
// rollout's policy, a 3-layer tanh MLP (obs_dim -> 128 -> 128 -> act_dim + 1) on the fused sensorPrev matrix.
//
// Separate torch kernels (two GEMMs, a head, two tanh passes, half a dozen elementwise kernels) cost the rollout
// ~0.09 ms per step -- a quarter of the simulator's own step -- and they are serial with it by data dependence.  Here:
//   * one persistent block per SM keeps the three weight matrices in shared memory (bfloat16, ~100 KB, rows padded by
//     16 bytes so that ldmatrix is conflict-free);
//   * every warp owns tiles of 16 observation rows end to end, so there is no block barrier after the prologue: the rows
//     arrive as float32 through cp.async (coalesced 16-byte chunks, no registers) into the warp's private staging slab,
//     are converted to bfloat16 A fragments on the fly, and the next tile's copy is issued as soon as layer 1 has read
//     the slab -- layer 2 and the head of this tile run while it is in flight, and the other warps of the SM cover the rest;
//   * the three layers run on the tensor cores (mma.sync m16n8k16, bfloat16 operands, float32 accumulators; B fragments by
//     ldmatrix.x4); the hidden activations never leave the registers -- the accumulator fragment of one layer, after bias +
//     tanh + bfloat16 rounding, IS the A fragment of the next;
//   * the epilogue writes the action (mid + half * tanh(mu + noise * scale)) straight into the row ftl_step consumes, plus
//     the value.
// Build: part of libftl.so (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ftl.h"

void ftl_set_error_message(const char* msg);

namespace {

constexpr int kHid = 128;          // hidden width of both layers
constexpr int kWarps = 8;          // warps per block (fewer when obs_dim is so large that eight slabs do not fit); each owns 16-row tiles
constexpr int kHidStride = kHid + 8;   // bf16 elements per shared-memory row of W2 / W3 (16 bytes of padding)

__device__ __forceinline__ void mma_bf16(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// shared-memory operands are 32-bit shared-window addresses, computed once per pointer
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void cp_async16(uint32_t a, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// hidden activations: one MUFU instruction (|error| ~ 2^-11, below the bfloat16 rounding the result goes through)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// rows x cols bfloat16 from global (row stride src_stride) into shared memory (row stride dst_stride), 16 bytes per chunk
__device__ __forceinline__ void copy_rows_async(__nv_bfloat16* dst, int dst_stride, const uint16_t* src, int src_stride, int rows,
                                                int cols, int tid, int nthreads) {
    const int per_row = cols / 8, total = rows * per_row;
    int r = tid / per_row, c = tid - r * per_row;           // one division per thread; then incremental
    const int dr = nthreads / per_row, dc = nthreads - dr * per_row;
    for (int e = tid; e < total; e += nthreads) {
        cp_async16(smem_addr(dst + (size_t)r * dst_stride + c * 8), src + (size_t)r * src_stride + c * 8);
        r += dr; c += dc;
        if (c >= per_row) { c -= per_row; r++; }
    }
}

// the warp's next tile of 16 observation rows -> its float32 staging slab [16][D + 8]: row by row, lane l copies the
// 16-byte chunks l, l + 32, ... of the row (coalesced); rows past the end of the batch repeat the last one
__device__ __forceinline__ void stage_rows_async(uint32_t slab_lane, int slab_stride_bytes, const float* obs, int obs_stride, int row0,
                                                 int n, int per_row, int lane) {
    const int last = n - 1 - row0;     // >= 0
    const float* src = obs + (size_t)row0 * obs_stride + lane * 4;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const float* s = src + (size_t)(r < last ? r : last) * obs_stride;
        const uint32_t d = slab_lane + r * slab_stride_bytes;
        for (int c = lane, o = 0; c < per_row; c += 32, o += 128) cp_async16(d + o * 4, s + o);
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(32 * kWarps, 1)
k_policy_mlp(const FtlMlpWeights w, const float* __restrict__ obs, const float* __restrict__ noise, int n,
             float* __restrict__ actions, float* __restrict__ values, int obs_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int D = w.obs_dim, xs = D + 8;
    __nv_bfloat16* W1 = reinterpret_cast<__nv_bfloat16*>(smem);                 // [128][xs]
    __nv_bfloat16* W2 = W1 + (size_t)kHid * xs;                                 // [128][136]
    __nv_bfloat16* W3 = W2 + (size_t)kHid * kHidStride;                         // [8][136]
    float* B1 = reinterpret_cast<float*>(W3 + (size_t)8 * kHidStride);          // [128], [128], [8]
    float* B2 = B1 + kHid;
    float* B3 = B2 + kHid;
    float* slabs = B3 + 8;                                                      // kWarps x [16][xs] float32
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n_out = w.act_dim + 1;
    float* slab = slabs + (size_t)warp * 16 * xs;
    const uint32_t slab_lane = smem_addr(slab + lane * 4);
    const int per_row = D / 4;
    // per-lane ldmatrix addresses of the B fragments: lanes 0-7 / 8-15 / 16-23 / 24-31 address the rows of the four 8x8
    // matrices (column tile nt at k0, nt at k0 + 8, nt + 1 at k0, nt + 1 at k0 + 8)
    const uint32_t w1_lane = smem_addr(W1 + (size_t)((lane >> 4) * 8 + (lane & 7)) * xs + ((lane >> 3) & 1) * 8);
    const uint32_t w2_lane = smem_addr(W2 + (size_t)((lane >> 4) * 8 + (lane & 7)) * kHidStride + ((lane >> 3) & 1) * 8);
    const uint32_t w3_lane = smem_addr(W3 + (size_t)(lane & 7) * kHidStride + ((lane >> 3) & 1) * 8);
    const uint32_t w1_pair = 16 * xs * 2;     // bytes between column-tile pairs of W1
    // epilogue constants of this lane's two output columns
    float e_scale[2], e_mid[2], e_half[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int c = 2 * t + j;
        e_scale[j] = (noise && c < w.act_dim) ? w.noise_scale[c] : 0.f;
        e_mid[j] = c < w.act_dim ? w.act_mid[c] : 0.f;
        e_half[j] = c < w.act_dim ? w.act_half[c] : 0.f;
    }
    const int n_tiles = (n + 15) / 16;
    const int n_warps = blockDim.x >> 5;
    int tile = blockIdx.x * n_warps + warp;
    const int tile_step = gridDim.x * n_warps;
    // ---- prologue: the warp's first tile and the weights, all asynchronous -------------------------------------------
    if (tile < n_tiles) stage_rows_async(slab_lane, xs * 4, obs, obs_stride, tile * 16, n, per_row, lane);
    copy_rows_async(W1, xs, w.w1, D, kHid, D, tid, blockDim.x);
    copy_rows_async(W2, kHidStride, w.w2, kHid, kHid, kHid, tid, blockDim.x);
    cp_async_commit();
    for (int e = tid; e < 8 * (kHid / 2); e += blockDim.x) {
        const int r = e / (kHid / 2), c = (e - r * (kHid / 2)) * 2;
        uint32_t v = 0;
        if (r < n_out) v = *reinterpret_cast<const uint32_t*>(w.w3 + (size_t)r * kHid + c);
        *reinterpret_cast<uint32_t*>(W3 + (size_t)r * kHidStride + c) = v;
    }
    for (int e = tid; e < kHid; e += blockDim.x) { B1[e] = w.b1[e]; B2[e] = w.b2[e]; }
    if (tid < 8) B3[tid] = tid < n_out ? w.b3[tid] : 0.f;
    cp_async_wait_all();
    __syncthreads();   // the only block barrier: the weights are in place
    for (; tile < n_tiles; tile += tile_step) {
        const int row0 = tile * 16;
        // the exploration noise of this lane's outputs: loaded now, used in the epilogue
        float e_noise[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        if (noise) {
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int row = row0 + g + 8 * half, c = 2 * t + j;
                    if (row < n && c < w.act_dim) e_noise[half][j] = noise[(size_t)row * w.act_dim + c];
                }
        }
        cp_async_wait_all();
        __syncwarp();      // every lane's chunks of this tile have landed
        // ---- layer 1: [16 x D] x [D x 128]; A fragments converted from the float32 slab -----------------------------
        float acc[kHid / 8][4];
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const float* x0 = slab + (size_t)g * xs + 2 * t;
        const float* x1 = x0 + (size_t)8 * xs;
        for (int k0 = 0; k0 < D; k0 += 16) {
            const float2 v0 = *reinterpret_cast<const float2*>(x0 + k0), v1 = *reinterpret_cast<const float2*>(x1 + k0);
            const float2 v2 = *reinterpret_cast<const float2*>(x0 + k0 + 8), v3 = *reinterpret_cast<const float2*>(x1 + k0 + 8);
            const uint32_t a0 = pack_bf16(v0.x, v0.y), a1 = pack_bf16(v1.x, v1.y), a2 = pack_bf16(v2.x, v2.y), a3 = pack_bf16(v3.x, v3.y);
            uint32_t wa = w1_lane + k0 * 2;
#pragma unroll
            for (int nt = 0; nt < kHid / 8; nt += 2, wa += w1_pair) {
                uint32_t b[4];
                ldmatrix_x4(b[0], b[1], b[2], b[3], wa);
                mma_bf16(acc[nt], a0, a1, a2, a3, b[0], b[1]);
                mma_bf16(acc[nt + 1], a0, a1, a2, a3, b[2], b[3]);
            }
        }
        __syncwarp();      // the slab has been read by every lane: the next tile may overwrite it
        if (tile + tile_step < n_tiles) stage_rows_async(slab_lane, xs * 4, obs, obs_stride, (tile + tile_step) * 16, n, per_row, lane);
        // bias + tanh + bfloat16: the accumulator fragments become layer 2's A fragments (k-step ks <- column tiles 2ks, 2ks+1)
        uint32_t h[kHid / 16][4];
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int nt = 2 * ks + half;
                const float2 bb = *reinterpret_cast<const float2*>(B1 + nt * 8 + 2 * t);
                h[ks][2 * half] = pack_bf16(tanh_fast(acc[nt][0] + bb.x), tanh_fast(acc[nt][1] + bb.y));       // row g
                h[ks][2 * half + 1] = pack_bf16(tanh_fast(acc[nt][2] + bb.x), tanh_fast(acc[nt][3] + bb.y));   // row g + 8
            }
        }
        // ---- layer 2: [16 x 128] x [128 x 128] ------------------------------------------------------------------------
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
#pragma unroll
            for (int nt = 0; nt < kHid / 8; nt += 2) {
                uint32_t b[4];
                ldmatrix_x4(b[0], b[1], b[2], b[3], w2_lane + (nt * 8 * kHidStride + ks * 16) * 2);
                mma_bf16(acc[nt], h[ks][0], h[ks][1], h[ks][2], h[ks][3], b[0], b[1]);
                mma_bf16(acc[nt + 1], h[ks][0], h[ks][1], h[ks][2], h[ks][3], b[2], b[3]);
            }
        }
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int nt = 2 * ks + half;
                const float2 bb = *reinterpret_cast<const float2*>(B2 + nt * 8 + 2 * t);
                h[ks][2 * half] = pack_bf16(tanh_fast(acc[nt][0] + bb.x), tanh_fast(acc[nt][1] + bb.y));
                h[ks][2 * half + 1] = pack_bf16(tanh_fast(acc[nt][2] + bb.x), tanh_fast(acc[nt][3] + bb.y));
            }
        }
        // ---- head: [16 x 128] x [128 x 8]; columns < act_dim are the action mean, column act_dim the value -------------
        float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
            uint32_t b0, b1;
            ldmatrix_x2(b0, b1, w3_lane + ks * 16 * 2);
            mma_bf16(out, h[ks][0], h[ks][1], h[ks][2], h[ks][3], b0, b1);
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int row = row0 + g + 8 * half;
            if (row >= n) continue;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = 2 * t + j;
                const float v = out[2 * half + j] + B3[c];
                if (c < w.act_dim) actions[(size_t)row * w.act_dim + c] = e_mid[j] + e_half[j] * tanhf(v + e_noise[half][j] * e_scale[j]);
                else if (c == w.act_dim) values[row] = v;
            }
        }
    }
    cp_async_wait_all();
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
    const size_t fixed = sizeof(__nv_bfloat16) * ((size_t)kHid * xs + (size_t)kHid * kHidStride + 8 * kHidStride) + sizeof(float) * (2 * kHid + 8);
    const size_t slab = sizeof(float) * (size_t)16 * xs;
    static int s_dev = -1, s_sms = 0, s_optin = 0;   // per-device attributes, looked up once (the rollout calls this every step)
    static size_t s_smem_set = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev != s_dev) {
        e = cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) { s_dev = dev; s_smem_set = 0; }
    }
    int warps = kWarps;
    if (e == cudaSuccess) {
        while (warps > 0 && fixed + warps * slab > (size_t)s_optin) warps--;
        if (warps < 1) {
            ftl_set_error_message("ftl_policy_mlp: obs_dim too large for the shared memory of this device");
            return FTL_ERR_INVALID;
        }
    }
    const size_t smem = fixed + warps * slab;
    if (e == cudaSuccess && smem > s_smem_set) {
        e = cudaFuncSetAttribute(k_policy_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) s_smem_set = smem;
    }
    if (e == cudaSuccess) {
        const int tiles = (n + 15) / 16, blocks = (tiles + warps - 1) / warps;
        k_policy_mlp<<<blocks < s_sms ? blocks : s_sms, 32 * warps, smem, (cudaStream_t)cuda_stream>>>(*w, obs_dev, noise_dev, n,
                                                                                                      actions_dev, values_dev, obs_stride);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str());
        return FTL_ERR_CUDA;
    }
    return FTL_OK;
}
