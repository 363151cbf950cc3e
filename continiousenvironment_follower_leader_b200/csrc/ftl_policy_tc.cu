// ftl_policy_tc.cu -- the rollout's policy MLP (obs_dim -> 128 -> 128 -> act_dim + 1, see ftl_policy.cu) on the 5th
// generation tensor cores: tcgen05.mma with the accumulators in tensor memory.
//
// mma.sync runs on sm_100's legacy HMMA path (measured in ftl_policy.cu: ~10.6 cycles of tensor pipe per HMMA.16816 and SM
// sub-partition, and every B fragment crosses shared memory -> registers once per 16 rows): 0.039 ms for 65 536 rows,
// whatever the number of warps.  Here one CTA per SM works on tiles of 128 rows:
//
//   loaders   (8 warps)  observation rows float32 -> bfloat16, written into shared memory in the K-major "interleaved"
//                        (no-swizzle) canonical layout of a tcgen05 operand: 8 x 16-byte core matrices, the two K halves
//                        of an MMA 128 bytes apart (LBO), 8-row groups (K / 8) * 128 bytes apart (SBO);
//   issuer    (1 lane)   layer 1 = obs_dim / 16 tcgen05.mma (M 128, N 128, K 16) of the X tile against W1 into TMEM
//                        columns [0, 128); layer 2 = 8 of the activations against W2 into the same columns; head = 8 with
//                        N = 16 into columns [128, 144).  tcgen05.commit signals the mbarriers;
//   epilogue  (16 warps) a thread owns 32 columns of a row of the tile (TMEM lane = row): tcgen05.ld 32 columns at a time, bias + tanh.approx
//                        + bfloat16 rounding, 16-byte stores into the activations' operand buffer (same canonical layout);
//                        after the head: action = mid + half * tanh(mu + noise * scale) and the value, to global memory.
//
// The weights are staged once per CTA with cp.async (a 16-byte chunk of a row of W is a row of a core matrix).  X is
// single-buffered: its loaders refill it as soon as layer 1 of the current tile has been committed, i.e. under the two
// epilogues and the remaining MMAs.  Everything that hands data between the generic proxy (st.shared) and the tensor core
// goes through fence.proxy.async + an mbarrier; TMEM reads are ordered with tcgen05.fence / tcgen05.wait::ld.
//
// Measured (B200, 65 536 x 240 -> 128 -> 128 -> 3): 33.5 us against 37.2 us for the mma.sync kernel.  A timeline of one CTA
// (-DFTL_POLICY_TRACE, globaltimer): 4.3 us prologue (weights), ~6 us until the first tile is staged (DRAM latency under
// the burst of 148 CTAs is ~3 us per batch of loads), then per tile layer 1 0.7 us -> epilogue 1.0-1.3 -> layer 2 0.3 ->
// epilogue 1.0 -> head 0.3 (the no-swizzle operand layout costs the MMAs ~3x their peak rate, which does not matter
// here), while the loaders need 4.4 us per tile: the second batch of a tile's loads is issued after X is free and its
// latency is exposed.  What bounds the kernel is bytes in flight: 6.5 TB/s x 3 us = 132 KB per SM, i.e. a whole float32 tile,
// and neither the registers of 8 loader warps (64 KB) nor the 36 KB of shared memory left beside weights, X and H hold
// it.  16 loader warps (a whole tile in flight) or 16 epilogue warps were slower (39.8 / 42.5 us): the extra warps take
// issue slots from the single thread that feeds the tensor core and from the epilogue.  The next step is the input as
// bfloat16 from the ray kernel (half the bytes, TMA straight into the operand layout, X double-buffered).
// Build: part of libftl.so (nvcc -gencode arch=compute_100a,code=sm_100a).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ftl.h"

void ftl_set_error_message(const char* msg);

#ifdef FTL_POLICY_TRACE   // diagnostic build: block 0 prints a timeline of its roles (globaltimer, ns since the kernel's start)
#include <stdio.h>
#define TR_DECL unsigned long long tr_t[40]; int tr_e[40]; int tr_n = 0; unsigned long long tr_0 = 0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_0));
#define TR(ev) do { if (blockIdx.x == 0 && lane == 0 && tr_n < 40) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tr_t[tr_n] = t_ - tr_0; tr_e[tr_n++] = (ev); } } while (0)
#define TR_DUMP(role) do { if (blockIdx.x == 0 && lane == 0) for (int i_ = 0; i_ < tr_n; i_++) printf("%s warp %d ev %d t %llu\n", role, warp, tr_e[i_], tr_t[i_]); } while (0)
#else
#define TR_DECL
#define TR(ev)
#define TR_DUMP(role)
#endif

namespace {

constexpr int kHid = 128;          // hidden width = N of the two big MMAs = rows of W1 / W2
constexpr int kRows = 128;         // rows per tile = M of every MMA = TMEM lanes
constexpr int kHeadN = 16;         // N of the head MMA (act_dim + 1 <= 8 real columns)
#ifndef FTL_POL_EPI
#define FTL_POL_EPI 8
#endif
#ifndef FTL_POL_LOAD
#define FTL_POL_LOAD 8
#endif
constexpr int kEpiWarps = FTL_POL_EPI, kLoadWarps = FTL_POL_LOAD;   // epilogue warp e: TMEM lanes 32 * (e % 4) .., columns kEpiCols * (e / 4) ..
constexpr int kEpiCols = kHid / (kEpiWarps / 4);             // hidden columns per epilogue thread (64 or 32)
constexpr int kIssuerWarp = kEpiWarps;
constexpr int kFirstLoadWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 1 + kLoadWarps);
constexpr int kLoadBatch = 8;      // items (8 rows x 32 columns) a loader warp has in flight: 16 float4 per lane
constexpr int kBatches = (kRows / 8) / kLoadWarps;           // row groups per loader warp and tile (2 or 1)
static_assert(kEpiWarps % 4 == 0 && kEpiCols % 32 == 0 && kBatches * kLoadWarps == kRows / 8 && kThreads <= 1024, "role layout");
constexpr int kTmemCols = 256;     // [0, 128) the layer accumulator, [128, 144) the head

// ---- small PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spins > (1 << 26)) __trap();   // a barrier that never completes is a bug, not a hang
    }
}
// the same for the roles that are not on the critical path: a waiter that spins steals issue slots from the one thread
// that feeds the tensor core (measured: 360 cycles per tcgen05.mma issue with 24 warps spinning beside it)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done) {
            __nanosleep(100);
            if (spins > (1 << 22)) __trap();
        }
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major; idesc: instruction descriptor; accumulate = 0 overwrites D
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive TMEM columns of this thread's lane (32 lanes x 32 bit per warp access)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Shared-memory matrix descriptor of a K-major, no-swizzle operand (cute::UMMA::SmemDescriptor): start address, leading
// byte offset (between the two 8-column halves of a K = 16 step) and stride byte offset (between 8-row groups), all in
// 16-byte units; bits 46-47 = 1 (Blackwell descriptor version); layout type 0 = no swizzle.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint64_t lo = (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16);
    const uint64_t hi = (uint64_t)(sbo_bytes >> 4) | (1ull << 14);
    return lo | (hi << 32);
}
// Instruction descriptor of kind::f16 (cute::UMMA::InstrDescriptor): D float32, A and B bfloat16, both K-major
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Element (row, k) of a K-major interleaved operand with K columns lives at byte
//   (row / 8) * (K / 8) * 128 + (k / 8) * 128 + (row % 8) * 16 + (k % 8) * 2        (8 x 8 core matrices of 128 bytes)

struct Smem {   // offsets into the dynamic shared memory, all multiples of 128 bytes
    uint32_t w1, w2, w3, x, h, bias, bars, tmem_slot, total;
};
__host__ __device__ inline Smem smem_layout(int D) {
    Smem s;
    uint32_t o = 0;
    s.w1 = o; o += kHid * D * 2;
    s.w2 = o; o += kHid * kHid * 2;
    s.w3 = o; o += kHeadN * kHid * 2;
    s.x = o; o += kRows * D * 2;
    s.h = o; o += kRows * kHid * 2;
    s.bias = o; o += (2 * kHid + kHeadN + 32) * 4;   // B1, B2, B3, then noise scale / action mid / action half (8 each)
    o = (o + 127u) & ~127u;
    s.bars = o; o += 8 * 8;          // x_full, x_empty, d_full, a_full, o_full
    s.tmem_slot = o; o += 16;
    s.total = (o + 127u) & ~127u;
    return s;
}

// rows x K bfloat16 of a torch.nn.Linear weight ([rows][K] row-major in global memory; rows >= real_rows are zero) into the
// operand layout: one 16-byte chunk = 8 consecutive k of one row = one row of a core matrix
__device__ __forceinline__ void stage_weight(uint32_t dst, const uint16_t* src, int rows, int real_rows, int K, int tid, int nthreads) {
    // thread -> (row of a core matrix r8, running core-matrix index m = row group * (K / 8) + k8): eight consecutive
    // threads fill the 128 contiguous bytes of one core matrix; no division in the loop
    const int cpr = K >> 3, n_mats = (rows >> 3) * cpr;
    const int r8 = tid & 7, step = nthreads >> 3;
    int m = tid >> 3, rg = m / cpr, k8 = m - rg * cpr;
    const int drg = step / cpr, dk8 = step - drg * cpr;
    for (; m < n_mats; m += step) {
        const int row = rg * 8 + r8;
        const uint32_t d = dst + (uint32_t)m * 128 + r8 * 16;
        if (row < real_rows) cp_async16(d, src + (size_t)row * K + k8 * 8);
        else sts128(d, 0u, 0u, 0u, 0u);
        rg += drg; k8 += dk8;
        if (k8 >= cpr) { k8 -= cpr; rg++; }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
k_policy_mlp_tc(const FtlMlpWeights w, const float* __restrict__ obs, const float* __restrict__ noise, int n,
                float* __restrict__ actions, float* __restrict__ values, int obs_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int D = w.obs_dim;
    const Smem L = smem_layout(D);
    const uint32_t base = smem_u32(smem);
    const uint32_t sW1 = base + L.w1, sW2 = base + L.w2, sW3 = base + L.w3, sX = base + L.x, sH = base + L.h;
    float* bias = reinterpret_cast<float*>(smem + L.bias);          // B1[128], B2[128], B3[16]
    const uint32_t bar_x_full = base + L.bars, bar_x_empty = bar_x_full + 8, bar_d_full = bar_x_full + 16, bar_a_full = bar_x_full + 24,
                   bar_o_full = bar_x_full + 32;   // the head's accumulator has a barrier of its own: a waiter must never fall two phases behind
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_out = w.act_dim + 1;
    const int n_tiles = (n + kRows - 1) / kRows;
    TR_DECL

    // ---- setup (before the wait for the kernel in front: weights and barriers do not depend on it) -------------------------
    asm volatile("griddepcontrol.launch_dependents;");   // ftl_step's k_kin (FTL_OPT_KIN_PDL) waits for this grid on the device
    if (warp == kIssuerWarp) {   // one warp allocates the tensor memory (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L.tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar_x_full, kLoadWarps);
        mbar_init(bar_x_empty, 1);
        mbar_init(bar_d_full, 1);
        mbar_init(bar_a_full, kEpiWarps);
        mbar_init(bar_o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp >= kFirstLoadWarp && (int)blockIdx.x < n_tiles) {   // the first tile's rows on their way into L2 while the weights are staged
        const int row = blockIdx.x * kRows + (warp - kFirstLoadWarp) * 16 + (lane >> 1);
        if (row < n) {
            const char* p = reinterpret_cast<const char*>(obs + (size_t)row * obs_stride) + (lane & 1) * 512;
            for (int o = 0; o < 512 && (lane & 1) * 512 + o < D * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        }
    }
    stage_weight(sW1, w.w1, kHid, kHid, D, tid, kThreads);
    stage_weight(sW2, w.w2, kHid, kHid, kHid, tid, kThreads);
    stage_weight(sW3, w.w3, kHeadN, n_out, kHid, tid, kThreads);
    for (int e = tid; e < kHid; e += kThreads) { bias[e] = w.b1[e]; bias[kHid + e] = w.b2[e]; }
    if (tid < kHeadN) bias[2 * kHid + tid] = tid < n_out ? w.b3[tid] : 0.f;
    if (tid < 8) {
        float* ec = bias + 2 * kHid + kHeadN;
        ec[tid] = (noise && tid < w.act_dim) ? w.noise_scale[tid] : 0.f;
        ec[8 + tid] = tid < w.act_dim ? w.act_mid[tid] : 0.f;
        ec[16 + tid] = tid < w.act_dim ? w.act_half[tid] : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_async_smem();          // the weights were written through the generic proxy; the tensor core reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the observations (and whatever else the kernel in front wrote) are complete
    TR(0);

    if (warp >= kFirstLoadWarp) {
        // ===== loaders: obs rows -> bfloat16 operand X ===================================================================
        // an item = 8 rows x 32 columns: lane (r = lane & 7, q = lane >> 3) converts 8 floats of row r; a warp's item is
        // 8 x 128 contiguous bytes of global memory and 4 core matrices = 512 contiguous bytes of shared memory
        // Batch b of a tile = row group lw + kLoadWarps * b (8 rows), all of its k-quads (u = 0 .. 7): every address is a per-lane base
        // plus a compile-time offset.  The loads of a tile's first batch are issued BEFORE the wait for X to become free
        // (they do not touch X); a second batch's latency is exposed (with 16 loader warps there is none: a whole tile is in
        // flight, but the extra warps slow the other roles down -- measured).
        const int lw = warp - kFirstLoadWarp;
        const int r8 = lane & 7, q = lane >> 3;
        float4 v[kLoadBatch][2];
        auto load_batch = [&](int tile, int b) {
            const int row = tile * kRows + (lw + kLoadWarps * b) * 8 + r8;
            const float* src = obs + (size_t)(row < n ? row : 0) * obs_stride + q * 8;
#pragma unroll
            for (int u = 0; u < kLoadBatch; u++) {
                v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u * 32 + q * 8 < D && row < n) {
                    v[u][0] = __ldg(reinterpret_cast<const float4*>(src + u * 32));
                    v[u][1] = __ldg(reinterpret_cast<const float4*>(src + u * 32 + 4));
                }
            }
        };
        auto store_batch = [&](int b) {
            const uint32_t dst = sX + (uint32_t)((lw + kLoadWarps * b) * (D >> 3) * 128 + q * 128 + r8 * 16);
#pragma unroll
            for (int u = 0; u < kLoadBatch; u++)
                if (u * 32 + q * 8 < D)
                    sts128(dst + u * 512, pack_bf16(v[u][0].x, v[u][0].y), pack_bf16(v[u][0].z, v[u][0].w),
                           pack_bf16(v[u][1].x, v[u][1].y), pack_bf16(v[u][1].z, v[u][1].w));
        };
        uint32_t ph_empty = 0;
        int tile = blockIdx.x;
        if (tile < n_tiles) load_batch(tile, 0);
        for (int it = 0; tile < n_tiles; it++) {
            if (it > 0) { mbar_wait_relaxed(bar_x_empty, ph_empty); ph_empty ^= 1; }   // layer 1 of the previous tile has read X
            TR(1);
            store_batch(0);
#pragma unroll
            for (int b = 1; b < kBatches; b++) { load_batch(tile, b); store_batch(b); }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_x_full);
            TR(2);
            tile += gridDim.x;
            if (tile < n_tiles) load_batch(tile, 0);
        }
        if (lw == 0) TR_DUMP("loader");
    } else if (warp == kIssuerWarp) {
        // ===== issuer: one lane feeds the tensor core =================================================================
        if (lane == 0) {
            const uint32_t idesc_hid = instr_desc(kRows, kHid), idesc_head = instr_desc(kRows, kHeadN);
            const uint32_t sbo_x = (uint32_t)(D >> 3) * 128, sbo_h = (kHid >> 3) * 128;
            uint32_t ph_full = 0, ph_a = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_x_full, ph_full); ph_full ^= 1;
                TR(10);
                tc_fence_after();
                for (int ks = 0; ks < (D >> 4); ks++)          // layer 1: X [128 x D] . W1^T -> TMEM columns [0, 128)
                    tc_mma(tmem, smem_desc(sX + ks * 256, 128, sbo_x), smem_desc(sW1 + ks * 256, 128, sbo_x), idesc_hid, ks > 0);
                tc_commit(bar_x_empty);                        // X may be refilled ...
                tc_commit(bar_d_full);                         // ... and the accumulator read
                TR(11);
                mbar_wait(bar_a_full, ph_a); ph_a ^= 1;        // the epilogue has left layer 1's activations in H
                TR(12);
                tc_fence_after();
                for (int ks = 0; ks < (kHid >> 4); ks++)       // layer 2: H . W2^T -> the same columns
                    tc_mma(tmem, smem_desc(sH + ks * 256, 128, sbo_h), smem_desc(sW2 + ks * 256, 128, sbo_h), idesc_hid, ks > 0);
                tc_commit(bar_d_full);
                TR(13);
                mbar_wait(bar_a_full, ph_a); ph_a ^= 1;
                TR(14);
                tc_fence_after();
                for (int ks = 0; ks < (kHid >> 4); ks++)       // head: H . W3^T (N = 16) -> columns [128, 144)
                    tc_mma(tmem + kHid, smem_desc(sH + ks * 256, 128, sbo_h), smem_desc(sW3 + ks * 256, 128, sbo_h), idesc_head, ks > 0);
                tc_commit(bar_o_full);
                TR(15);
            }
            TR_DUMP("issuer");
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp e handles TMEM lanes 32 * (e % 4) .. (a warp can only read its own quarter of the lanes) and
        //       the columns 32 * (e / 4) .. of the two hidden layers; the head's 8 columns are read by warps 0..3 ==========
        const int quarter = warp & 3, cgroup = warp >> 2;   // columns [kEpiCols * cgroup, + kEpiCols)
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
        const uint32_t h_row = sH + (uint32_t)((row_in_tile >> 3) * (kHid >> 3) * 128 + (row_in_tile & 7) * 16);
        const float* ec = bias + 2 * kHid + kHeadN;   // noise scale [8], action mid [8], action half [8]
        uint32_t ph_d = 0, ph_o = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int row = tile * kRows + row_in_tile;
            float nz[8];   // raw noise of this row: loaded now, first used after the head
#pragma unroll
            for (int c = 0; c < 8; c++)
                nz[c] = (cgroup == 0 && noise && row < n && c < w.act_dim) ? noise[(size_t)row * w.act_dim + c] : 0.f;
#pragma unroll 1
            for (int layer = 0; layer < 2; layer++) {
                mbar_wait(bar_d_full, ph_d); ph_d ^= 1;
                TR(20 + layer);
                tc_fence_after();
                const float* b = bias + layer * kHid;
#pragma unroll
                for (int cc = 0; cc < kEpiCols; cc += 32) {
                    const int c0 = cgroup * kEpiCols + cc;
                    uint32_t v[32];
                    tc_ld32(t_lane + c0, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; j++) {      // 8 columns = one 16-byte row of a core matrix
                        const float4 b0 = *reinterpret_cast<const float4*>(b + c0 + j * 8), b1 = *reinterpret_cast<const float4*>(b + c0 + j * 8 + 4);
                        const uint32_t p0 = pack_bf16(tanh_fast(__uint_as_float(v[j * 8 + 0]) + b0.x), tanh_fast(__uint_as_float(v[j * 8 + 1]) + b0.y));
                        const uint32_t p1 = pack_bf16(tanh_fast(__uint_as_float(v[j * 8 + 2]) + b0.z), tanh_fast(__uint_as_float(v[j * 8 + 3]) + b0.w));
                        const uint32_t p2 = pack_bf16(tanh_fast(__uint_as_float(v[j * 8 + 4]) + b1.x), tanh_fast(__uint_as_float(v[j * 8 + 5]) + b1.y));
                        const uint32_t p3 = pack_bf16(tanh_fast(__uint_as_float(v[j * 8 + 6]) + b1.z), tanh_fast(__uint_as_float(v[j * 8 + 7]) + b1.w));
                        sts128(h_row + (uint32_t)((c0 >> 3) + j) * 128, p0, p1, p2, p3);
                    }
                }
                tc_fence_before();       // this thread's TMEM reads are done before the issuer overwrites the columns
                fence_async_smem();      // H was written through the generic proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_full);
                TR(22 + layer);
            }
            if (cgroup == 0) {
                mbar_wait(bar_o_full, ph_o); ph_o ^= 1;
                TR(24);
                tc_fence_after();
                uint32_t o[8];
                tc_ld8(t_lane + kHid, o);
                tc_wait_ld();
                tc_fence_before();
                if (row < n) {
                    const float* b3 = bias + 2 * kHid;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float vv = __uint_as_float(o[c]) + b3[c];
                        if (c < w.act_dim) actions[(size_t)row * w.act_dim + c] = ec[8 + c] + ec[16 + c] * tanhf(vv + nz[c] * ec[c]);
                        else if (c == w.act_dim) values[row] = vv;
                    }
                }
            }
        }
    }
    if (warp == 0) { TR(25); TR_DUMP("epilogue"); }
    // ---- teardown ----------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace

// 0: launched; 1: this shape is not for this kernel (the caller falls back to the mma.sync kernel); < 0: error
int ftl_policy_mlp_tc_launch(const FtlMlpWeights* w, const float* obs_dev, int32_t obs_stride, const float* noise_dev, int32_t n,
                             float* actions_dev, float* values_dev, cudaStream_t stream) {
    static int s_dev = -1, s_sms = 0, s_optin = 0;
    static size_t s_smem_set = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev != s_dev) {
        e = cudaDeviceGetAttribute(&s_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) { s_dev = dev; s_smem_set = 0; }
    }
    if (e != cudaSuccess) { ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str()); return FTL_ERR_CUDA; }
    const Smem L = smem_layout(w->obs_dim);
    if (w->obs_dim % 16 != 0 || w->obs_dim > 256 || (size_t)L.total > (size_t)s_optin || (obs_stride % 4) != 0) return 1;
    if ((size_t)L.total > s_smem_set) {
        e = cudaFuncSetAttribute(k_policy_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        if (e != cudaSuccess) { ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str()); return FTL_ERR_CUDA; }
        s_smem_set = L.total;
    }
    const int tiles = (n + kRows - 1) / kRows;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(tiles < s_sms ? tiles : s_sms); lc.blockDim = dim3(kThreads); lc.dynamicSmemBytes = L.total; lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // scheduled while the kernel in front drains; waits on the device
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    e = cudaLaunchKernelEx(&lc, k_policy_mlp_tc, *w, obs_dev, noise_dev, n, actions_dev, values_dev, obs_stride);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { ftl_set_error_message((std::string("ftl_policy_mlp: ") + cudaGetErrorString(e)).c_str()); return FTL_ERR_CUDA; }
    return 0;
}
