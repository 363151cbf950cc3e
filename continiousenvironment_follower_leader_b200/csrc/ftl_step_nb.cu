// ftl_step_nb.cu -- the kinematics / reset kernels for ONE value of NB (number of dynamic obstacles),
// compiled once per NB in 0..FTL_MAX_BEARS with -DFTL_NB=<n> (build.py) so the robots live in registers.
//
//   k_kin<NB>    one thread per env: action decode, the F sub-frames of robot kinematics fused in registers (follower,
//                waypoints, bears, leader, integer collision tests), per-frame records for the bookkeeping kernel, the
//                two tracker scans, the history snapshot, the observation outputs.  This is the part of Game.step the
//                ray kernel has to wait for.
//   k_reset<NB>  masked re-initialisation from the scenario pool
// The bookkeeping (k_book) and finishing (k_finish) kernels do not depend on NB and live in ftl_capi.cu.
// (A variant with one warp per robot role -- follower / leader / bears of 32 envs in three warps exchanging poses through
// shared memory, one block barrier per frame -- was built and measured this round: a third of the dependent chain per
// warp, but three times the threads no longer fit the register file in one wave (96 registers without spills, 2 048
// blocks of 96 threads), and two waves of 0.05 ms lose to one wave of 0.088 ms; profiles/r02_ab_log.txt.)
#include <cuda_runtime.h>

#include "ftl_launch.h"

#ifndef FTL_SYNC_EVERY
#define FTL_SYNC_EVERY 0
#endif
#if FTL_SYNC_EVERY > 0
__host__ __device__ static __forceinline__ void ftl_frame_sync(int f) {
#if defined(__CUDA_ARCH__)
    if (f % FTL_SYNC_EVERY == 0) __syncthreads();
#endif
}
#define FTL_FRAME_SYNC(f) ftl_frame_sync(f)
#endif
#include "ftl_step.cuh"
#include "ftl_reset_image.cuh"
#include "ftl_book.cuh"

#ifndef FTL_NB
#error "compile with -DFTL_NB=<number of bears>"
#endif

using namespace ftl;

#ifndef FTL_FUSED_BOOK
#define FTL_FUSED_BOOK 1   // see ftl_capi.cu
#endif
#ifndef FTL_STEP_THREADS
#define FTL_STEP_THREADS 64   // 65536 envs are only ~14 warps per SM: small blocks spread them evenly over the 148 SMs
#endif
#ifndef FTL_STEP_MINBLOCKS
#define FTL_STEP_MINBLOCKS 8    // caps the kernel at 128 registers: all ~14 warps/SM of a 65536-env batch resident in one wave
#endif
template <int NB>
__global__ void __launch_bounds__(FTL_STEP_THREADS, FTL_STEP_MINBLOCKS)
k_kin(const __grid_constant__ DevCfg cfg, const __grid_constant__ DevState s, const __grid_constant__ DevPool pool,
      const __grid_constant__ DevState img, const __grid_constant__ DevOutputs img_out,
      const void* __restrict__ actions, const __grid_constant__ DevOutputs out, double* __restrict__ stats, int seq,
      int fused_book, int rec_in_smem, int kin_pdl) {
    // k_book, k_rays and k_finish are launched as programmatic dependents of this kernel and of each other: their blocks
    // may start as soon as every block of the kernel in front is running, take the SM resources that finished blocks
    // free, and wait per group of 32 envs on kin_flag / book_flag.
    // Optionally (FTL_OPT_KIN_PDL, ftl_set_option) this kernel is itself a programmatic dependent of whatever is in front
    // of it in the stream: its blocks are scheduled while that kernel drains and wait here until it -- and everything
    // before it -- has completed and flushed.  Off by default: behind the previous step's k_finish the early blocks take
    // what the last ray blocks would have used (+2.5 % per step), behind the rollout's policy kernel +1.6 %
    // (profiles/r02_ab_log.txt (14), (17)).
    if (kin_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;   // whole warps: n is padded to a multiple of 32
    // envs that finished in the previous step start their next episode here (their first observation was already
    // delivered with that step's outputs): a copy of the scenario's reset image, made by the whole warp
    apply_pending_resets(cfg, s, pool, img, i);
    World<NB> w;
    Episode e;
    world_load<NB>(s, i, w);
    episode_load_kin(s, i, e);
    double a0, a1;
    decode_action(cfg.c, actions, i, s.n_real, &a0, &a1);
    KinCtx k;
    kin_begin<NB>(cfg, pool, a0, a1, w, e, k);
    // per-frame records: in shared memory when the bookkeeping runs in this thread (a record is read back ~0.05 ms after
    // it was written: from HBM that is a DRAM round trip per frame on the bookkeeping's dependent chain), in HBM for k_book
    extern __shared__ __align__(16) unsigned char rec_smem[];
    FrameRec rec = {s.rec_f + i, s.rec_l + i, s.rec_bits + i, s.rec_lbits + i, (size_t)s.n};
#if FTL_FUSED_BOOK && !defined(FTL_REC_GLOBAL)
    if (rec_in_smem) {
        const int T = blockDim.x, F = cfg.c.frames_per_step;
        float2* rf = reinterpret_cast<float2*>(rec_smem);
        float2* rl = rf + (size_t)F * T;
        unsigned char* rb = reinterpret_cast<unsigned char*>(rl + (size_t)F * T);
        rec = FrameRec{rf + threadIdx.x, rl + threadIdx.x, rb + threadIdx.x, nullptr, (size_t)T};
    }
#endif
    const int fps = env_frames(cfg, s, i);
    kin_frames<NB>(cfg, pool, i, w, e, k, rec, e.step_count, 0, fps, fps, env_draws(cfg, s, i));
    Tracker t;
    int snap_pushes;
    tracker_load(s, i, t, &snap_pushes);
    sense_serial<NB>(cfg, s, i, w, t, &snap_pushes, &e.overflow);
    tracker_store(s, i, t, snap_pushes);
    write_outputs_obs<NB>(cfg, pool, out, i, w, e.scenario, e.cur_target_id);
    world_store<NB>(s, i, w);
    episode_store_kin(s, i, e);
    // publish: every lane's stores, then the warp's flag (release); k_book and k_rays acquire it before reading this
    // group's state
    publish_group_flag(s.kin_flag, i >> 5, seq);
#if FTL_FUSED_BOOK   // the bookkeeping in the same thread instead of k_book (build option, A/B in profiles/r02_ab_log.txt)
    book_env(cfg, s, img_out, out, stats, pool.n_scenarios, i, &rec);
    publish_group_flag(s.book_flag, i >> 5, seq);
#else
    (void)fused_book; (void)stats; (void)rec_in_smem;
#endif
}

template <int NB>
__global__ void __launch_bounds__(128)
k_reset(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, const uint8_t* __restrict__ mask,
        const int* __restrict__ scenario_ids, const DevOutputs out, int reset_filler) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    if (i < s.n_real ? (mask && !mask[i]) : !reset_filler) return;
    World<NB> w;
    Episode e;
    int episodes = s.gi[(size_t)GI_EPISODE * s.n + i];
    int scen = (scenario_ids && i < s.n_real) ? scenario_ids[i] : next_scenario(cfg, pool.n_scenarios, i, episodes);
    scen = scen < 0 ? 0 : scen >= pool.n_scenarios ? pool.n_scenarios - 1 : scen;
    env_reset<NB>(cfg, s, pool, i, scen, w, e);
    write_outputs<NB>(cfg, pool, out, i, w, e, false);
    world_store<NB>(s, i, w);
    episode_store(s, i, e);
}


#define FTL_CAT2(a, b) a##b
#define FTL_CAT(a, b) FTL_CAT2(a, b)

void FTL_CAT(ftl_launch_kin_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const DevState& img,
                                         const DevOutputs& img_out, const void* actions, const DevOutputs& out,
                                         double* stats, int seq, int fused_book, int pdl, cudaStream_t st) {
#ifdef FTL_STEP_LAUNCH_THREADS
    int threads = FTL_STEP_LAUNCH_THREADS;
#else
    int threads = FTL_STEP_THREADS;
#endif
    int blocks = (s.n + threads - 1) / threads;
    // 17 bytes of record per frame and env; shared memory while that stays small (the ray blocks that move in beside
    // this kernel's last blocks need theirs), HBM otherwise
    size_t rec_bytes = (size_t)17 * cfg.c.frames_per_step * threads;
    const int rec_in_smem = (FTL_FUSED_BOOK && fused_book && rec_bytes <= 16 * 1024) ? 1 : 0;
    const int kin_pdl = pdl ? 1 : 0;   // FTL_OPT_KIN_PDL and neither per-kernel timing nor a stream that is being captured
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(blocks); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = rec_in_smem ? rec_bytes : 0; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = kin_pdl ? 1 : 0;
    cudaLaunchKernelEx(&lc, k_kin<FTL_NB>, cfg, s, pool, img, img_out, actions, out, stats, seq, fused_book, rec_in_smem, kin_pdl);
}
void FTL_CAT(ftl_launch_reset_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const uint8_t* mask,
                                           const int* ids, const DevOutputs& out, int reset_filler, cudaStream_t st) {
    int threads = 128, blocks = (s.n + threads - 1) / threads;
    k_reset<FTL_NB><<<blocks, threads, 0, st>>>(cfg, s, pool, mask, ids, out, reset_filler);
}
