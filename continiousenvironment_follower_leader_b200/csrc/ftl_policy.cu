// ftl_policy.cu -- a consumer of the simulator's observations: one fused kernel for the
// rollout's policy, a 3-layer tanh MLP (obs_dim -> 128 -> 128 -> act_dim + 1) on the fused sensorPrev matrix.
//
// Separate torch kernels (two GEMMs, a head, two tanh passes, half a dozen elementwise kernels) cost the rollout
// ~0.09 ms per step -- a quarter of the simulator's own step -- and they are serial with it by data dependence.  Here:
//   * one persistent block per SM keeps the three weight matrices in shared memory (bfloat16, ~100 KB, rows padded by
//     16 bytes so that ldmatrix is conflict-free);
//   * every warp owns tiles of 16 observation rows end to end, so there is no block barrier after the prologue.  The
//     rows are read as float4 (coalesced, eight loads in flight per lane), rounded to bfloat16 and parked in the warp's
//     private slab [16][max(obs_dim, 128) + 8]; at 7.9 KB per slab SIXTEEN warps fit beside the weights (a float32 slab filled by
//     cp.async allowed eight: the kernel was latency-bound at two warps per scheduler, 0.039 ms), and the other warps'
//     tensor-core work covers a warp's loads;
//   * the three layers run on the tensor cores (mma.sync m16n8k16, bfloat16 operands, float32 accumulators; A and B
//     fragments by ldmatrix.x4); a layer's activations (bias + tanh.approx + bfloat16 rounding) go back into the slab,
//     which is the next layer's A operand;
//   * the epilogue writes the action (mid + half * tanh(mu + noise * scale)) straight into the row ftl_step consumes, plus
//     the value.
// mma.sync runs on the legacy HMMA path of sm_100 (measured: ~10.6 cycles of tensor pipe per HMMA.16816 and SM
// sub-partition, a floor of 14 us for this shape; 37-39 us whatever the number of warps).  ftl_policy_tc.cu is the same
// network on tcgen05 / tensor memory (33.5 us) and is what ftl_policy_mlp launches whenever the shape fits it
// (obs_dim <= 256); this kernel serves the rest (and FTL_POLICY_IMPL=mma).
// Build: part of libftl.so (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "../../include/ftl.h"

void ftl_set_error_message(const char* msg);
int ftl_policy_mlp_tc_launch(const FtlMlpWeights* w, const float* obs_dev, int32_t obs_stride, const float* noise_dev, int32_t n,
                             float* actions_dev, float* values_dev, cudaStream_t stream);   // ftl_policy_tc.cu

namespace {

constexpr int kHid = 128;          // hidden width of both layers
constexpr int kWarps = 16;         // warps per block (fewer when obs_dim is so large that the slabs do not fit); each owns 16-row tiles
constexpr int kHidStride = kHid + 8;   // bf16 elements per shared-memory row of W2 / W3 (16 bytes of padding)
constexpr int kStageBatch = 8;     // float4 loads in flight per lane while a tile is staged

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
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t v0, uint32_t v1) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v0), "r"(v1) : "memory");
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

// a tile of 16 observation rows -> the warp's bfloat16 slab [16][xs]: the tile is walked as 16 * per_row chunks of four
// floats (chunk q = row q / per_row, column 4 * (q % per_row)), lane l takes chunks l, l + 32, ...: consecutive lanes read
// consecutive 16 bytes.  Rows past the end of the batch repeat the last one (their results are never written).
__device__ __forceinline__ void stage_rows(uint32_t slab, int xs, const float* obs, int obs_stride, int row0, int n, int per_row,
                                           int lane) {
    const int last = n - 1 - row0;     // >= 0
    const int total = 16 * per_row;
    int r = 0, c = lane;               // chunk `lane`
    while (c >= per_row) { c -= per_row; r++; }
    const int dr = 32 / per_row, dc = 32 - dr * per_row;
    for (int q0 = lane; q0 < total; q0 += 32 * kStageBatch) {
        float4 v[kStageBatch];
        uint32_t dst[kStageBatch];
#pragma unroll
        for (int u = 0; u < kStageBatch; u++) {
            const bool live = q0 + 32 * u < total;
            const int rr = r < last ? r : last;
            dst[u] = live ? slab + (uint32_t)(r * xs + c * 4) * 2 : 0xffffffffu;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) v[u] = __ldg(reinterpret_cast<const float4*>(obs + (size_t)(row0 + rr) * obs_stride + c * 4));
            r += dr; c += dc;
            if (c >= per_row) { c -= per_row; r++; }
        }
#pragma unroll
        for (int u = 0; u < kStageBatch; u++)
            if (dst[u] != 0xffffffffu) sts64(dst[u], pack_bf16(v[u].x, v[u].y), pack_bf16(v[u].z, v[u].w));
    }
}

__global__ void __launch_bounds__(32 * kWarps, 1)
k_policy_mlp(const FtlMlpWeights w, const float* __restrict__ obs, const float* __restrict__ noise, int n,
             float* __restrict__ actions, float* __restrict__ values, int obs_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int D = w.obs_dim, xs = D + 8;
    const int ss = (D > kHid ? D : kHid) + 8;   // slab row stride: the slab also holds the 128 hidden activations of a row
    __nv_bfloat16* W1 = reinterpret_cast<__nv_bfloat16*>(smem);                 // [128][xs]
    __nv_bfloat16* W2 = W1 + (size_t)kHid * xs;                                 // [128][136]
    __nv_bfloat16* W3 = W2 + (size_t)kHid * kHidStride;                         // [8][136]
    float* B1 = reinterpret_cast<float*>(W3 + (size_t)8 * kHidStride);          // [128], [128], [8]
    float* B2 = B1 + kHid;
    float* B3 = B2 + kHid;
    __nv_bfloat16* slabs = reinterpret_cast<__nv_bfloat16*>(B3 + 8);            // n_warps x [16][ss] bfloat16
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n_out = w.act_dim + 1;
    const int n_warps = blockDim.x >> 5;
    const uint32_t slab = smem_addr(slabs + (size_t)warp * 16 * ss);
    const int per_row = D / 4;
    // per-lane ldmatrix addresses.  B fragments: lanes 0-7 / 8-15 / 16-23 / 24-31 address the rows of the four 8x8 matrices
    // (column tile nt at k0, nt at k0 + 8, nt + 1 at k0, nt + 1 at k0 + 8).  A fragments: rows 0-7 at k0, rows 8-15 at k0,
    // rows 0-7 at k0 + 8, rows 8-15 at k0 + 8 (= a0, a1, a2, a3 of mma.m16n8k16).
    const uint32_t w1_lane = smem_addr(W1 + (size_t)((lane >> 4) * 8 + (lane & 7)) * xs + ((lane >> 3) & 1) * 8);
    const uint32_t w2_lane = smem_addr(W2 + (size_t)((lane >> 4) * 8 + (lane & 7)) * kHidStride + ((lane >> 3) & 1) * 8);
    const uint32_t w3_lane = smem_addr(W3 + (size_t)(lane & 7) * kHidStride + ((lane >> 3) & 1) * 8);
    const uint32_t a_lane = slab + (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * ss + (lane >> 4) * 8) * 2;
    const uint32_t h_lane = slab + (uint32_t)(g * ss + 2 * t) * 2;   // where this lane's accumulator columns go (row g; row g + 8: + 8 * xs)
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
    asm volatile("griddepcontrol.launch_dependents;");   // ftl_step's k_kin (FTL_OPT_KIN_PDL) waits for this grid on the device
    // tiles are dealt round-robin to the blocks and, inside a block, to its warps: every SM gets the same number (+-1)
    // ---- prologue: the weights, asynchronously, while every warp stages its first tile ----------------------------------
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
    int tile = blockIdx.x + gridDim.x * warp;
    const int tile_step = gridDim.x * n_warps;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent: the observations are complete from here on
    if (tile < n_tiles) stage_rows(slab, ss, obs, obs_stride, tile * 16, n, per_row, lane);
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
        __syncwarp();      // every lane's part of the tile is in the slab
        // ---- layer 1: [16 x D] x [D x 128] -----------------------------------------------------------------------------
        float acc[kHid / 8][4];
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        for (int k0 = 0; k0 < D; k0 += 16) {
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(a0, a1, a2, a3, a_lane + k0 * 2);
            uint32_t wa = w1_lane + k0 * 2;
#pragma unroll
            for (int nt = 0; nt < kHid / 8; nt += 2, wa += w1_pair) {
                uint32_t b[4];
                ldmatrix_x4(b[0], b[1], b[2], b[3], wa);
                mma_bf16(acc[nt], a0, a1, a2, a3, b[0], b[1]);
                mma_bf16(acc[nt + 1], a0, a1, a2, a3, b[2], b[3]);
            }
        }
        __syncwarp();      // the slab has been read by every lane: the activations may overwrite it
        // bias + tanh + bfloat16 -> the slab's first 128 columns: layer 2's A operand
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) {
            const float2 bb = *reinterpret_cast<const float2*>(B1 + nt * 8 + 2 * t);
            sts32(h_lane + nt * 16, pack_bf16(tanh_fast(acc[nt][0] + bb.x), tanh_fast(acc[nt][1] + bb.y)));                // row g
            sts32(h_lane + nt * 16 + 8 * ss * 2, pack_bf16(tanh_fast(acc[nt][2] + bb.x), tanh_fast(acc[nt][3] + bb.y)));   // row g + 8
        }
        __syncwarp();
        // ---- layer 2: [16 x 128] x [128 x 128] ------------------------------------------------------------------------
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(a0, a1, a2, a3, a_lane + ks * 32);
#pragma unroll
            for (int nt = 0; nt < kHid / 8; nt += 2) {
                uint32_t b[4];
                ldmatrix_x4(b[0], b[1], b[2], b[3], w2_lane + (nt * 8 * kHidStride + ks * 16) * 2);
                mma_bf16(acc[nt], a0, a1, a2, a3, b[0], b[1]);
                mma_bf16(acc[nt + 1], a0, a1, a2, a3, b[2], b[3]);
            }
        }
        __syncwarp();      // every lane has read layer 1's activations
#pragma unroll
        for (int nt = 0; nt < kHid / 8; nt++) {
            const float2 bb = *reinterpret_cast<const float2*>(B2 + nt * 8 + 2 * t);
            sts32(h_lane + nt * 16, pack_bf16(tanh_fast(acc[nt][0] + bb.x), tanh_fast(acc[nt][1] + bb.y)));
            sts32(h_lane + nt * 16 + 8 * ss * 2, pack_bf16(tanh_fast(acc[nt][2] + bb.x), tanh_fast(acc[nt][3] + bb.y)));
        }
        __syncwarp();
        // ---- head: [16 x 128] x [128 x 8]; columns < act_dim are the action mean, column act_dim the value -------------
        float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ks++) {
            uint32_t a0, a1, a2, a3, b0, b1;
            ldmatrix_x4(a0, a1, a2, a3, a_lane + ks * 32);
            ldmatrix_x2(b0, b1, w3_lane + ks * 16 * 2);
            mma_bf16(out, a0, a1, a2, a3, b0, b1);
        }
        __syncwarp();      // the slab is free for the next tile
        if (tile + tile_step < n_tiles) stage_rows(slab, ss, obs, obs_stride, (tile + tile_step) * 16, n, per_row, lane);
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
    {   // the tcgen05 kernel (ftl_policy_tc.cu) takes every shape it has room for; FTL_POLICY_IMPL=mma keeps the mma.sync one
        const char* impl = getenv("FTL_POLICY_IMPL");
        if (!(impl && impl[0] == 'm')) {
            const int rc = ftl_policy_mlp_tc_launch(w, obs_dev, obs_stride, noise_dev, n, actions_dev, values_dev, (cudaStream_t)cuda_stream);
            if (rc <= 0) return rc;
        }
    }
    const int xs = w->obs_dim + 8;
    const size_t fixed = sizeof(__nv_bfloat16) * ((size_t)kHid * xs + (size_t)kHid * kHidStride + 8 * kHidStride) + sizeof(float) * (2 * kHid + 8);
    const size_t slab = sizeof(__nv_bfloat16) * (size_t)16 * ((w->obs_dim > kHid ? w->obs_dim : kHid) + 8);
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
        const int tiles = (n + 15) / 16;
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(tiles < s_sms ? tiles : s_sms); lc.blockDim = dim3(32 * warps); lc.dynamicSmemBytes = smem;
        lc.stream = (cudaStream_t)cuda_stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        e = cudaLaunchKernelEx(&lc, k_policy_mlp, *w, obs_dev, noise_dev, n, actions_dev, values_dev, obs_stride);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str());
        return FTL_ERR_CUDA;
    }
    return FTL_OK;
}
