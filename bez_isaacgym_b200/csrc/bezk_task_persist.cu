// Persistent, software-pipelined form of the fused BezKick post-physics step.
//
// The one-shot tile kernel (bezk_task.cu) is latency-bound at 262 144 envs: every CTA loads, computes and stores once, its sparse
// gathers sit in registers while the warp waits, 16 warps per SM are resident and DRAM is busy 60 % of the time.  Here ONE CTA per
// SM stays resident; each of its PK_C warps owns a private shared-memory stage (one 32-env tile of every input) plus an output
// buffer, and walks its tiles with the loads of tile i+1 in flight while it computes tile i:
//
//   issue(tile)   * 6-8 cp.async.bulk (TMA 1-D) copies: the dense dof_state / root_states sub-tiles and the small per-env arrays
//                   (goal, ball_init, reset_buf, progress_buf, prev_lin_vel, values) -- all contiguous per tile;
//                 * per lane (= env) 5-7 cp.async.cg (LDGSTS, 16 bytes, .L2::64B, L1 bypassed) copies of the SPARSE rows: the
//                   16-byte aligned windows around the 40-byte IMU-link slice of rigid_body and the two 12-byte foot rows;
//                 everything completes on the warp's mbarrier (expect_tx for the bulk bytes, cp.async.mbarrier.arrive.noinc for
//                 the gathers): no register is held for data in flight.
//   loop          wait(full) -> masked resets patched in shared memory -> this env's rows into registers -> issue(next tile) into
//                 the SAME stage -> observation + reward from registers -> 54-float rows into the output buffer -> one bulk store;
//                 scalar outputs go straight to global memory (coalesced 128 / 256-byte runs per warp).
//
// (A first version with a dedicated producer warp feeding a ring in tile order lost a third of the consumers' time to
// head-of-line blocking on the "empty" barriers -- ncu stall samples, profiles/r02_persist.md; warps that feed themselves cannot
// block each other.)
// Results are bit-identical to task_tile_kernel<7, false, *, BEZK_TASK_KICK> (same device functions, same order of operations).
// Used for BezKick, no cleats, full 32-env tiles, 16-byte aligned dense tensors and 8-byte aligned sparse rows; every other case
// keeps the one-shot kernel.
#include "bezk_common.cuh"
#include "bezk_internal.h"
#include "bezk_task_math.cuh"

namespace bezk {

#ifndef BEZK_PK_WARPS
#define BEZK_PK_WARPS 10
#endif
constexpr int PK_C = BEZK_PK_WARPS;
constexpr int PK_THREADS = 32 * PK_C;
constexpr int PK_WT = 32;
// stage layout in bytes (all offsets multiples of 16)
constexpr int PK_DOF = 0;                          // 32 x 36 floats
constexpr int PK_ROOT = PK_DOF + PK_WT * 144;      // 32 x 26 floats
constexpr int PK_GOAL = PK_ROOT + PK_WT * 104;     // 32 x float2
constexpr int PK_BINIT = PK_GOAL + PK_WT * 8;
constexpr int PK_RESET = PK_BINIT + PK_WT * 8;     // 32 x int64
constexpr int PK_PROG = PK_RESET + PK_WT * 8;
constexpr int PK_PREV = PK_PROG + PK_WT * 8;       // 32 x 3 floats
constexpr int PK_VALUE = PK_PREV + PK_WT * 12;     // 32 floats
// Sparse rows arrive as 16-byte cp.async.cg chunks (L1 bypassed: with ~200 KB of shared memory the L1 is down to 28 KB, and
// L1-allocating 4 / 8-byte copies collapse -- measured).  The 16-byte aligned window around an 8-byte aligned slice starts at the
// slice or 8 bytes before it, so the data sits at +0 or +8 of the row (source address parity).
constexpr int PK_IMU = PK_VALUE + PK_WT * 4;       // 32 rows of 48 bytes: window of 3 chunks around the 40-byte slice
constexpr int PK_IMU_ROW = 48;
constexpr int PK_FEET = PK_IMU + PK_WT * PK_IMU_ROW;   // 32 rows of 64 bytes: left window (1-2 chunks) at +0, right window at +32
constexpr int PK_FEET_ROW = 64;
constexpr int PK_STAGE_BYTES = PK_FEET + PK_WT * PK_FEET_ROW;
static_assert(PK_STAGE_BYTES % 128 == 0, "stages stay 128-byte aligned");
constexpr int PK_OUT_BYTES = PK_WT * 54 * 4;       // one consumer's observation rows
constexpr int PK_WARP_BYTES = PK_STAGE_BYTES + PK_OUT_BYTES;      // one warp's private stage + output buffer
constexpr size_t PK_SMEM = (size_t)PK_C * PK_WARP_BYTES;
static_assert(PK_SMEM <= 227 * 1024, "too many warps for one SM's shared memory");

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
// the executing thread's earlier cp.async copies arrive on the mbarrier when they land (.noinc: counted in the init count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Issue every load of one 32-env tile into the warp's stage; all of them complete on `bar` (init count 1 + 32).
__device__ __forceinline__ void pk_issue(const TaskArgs& a, unsigned char* sb, uint64_t* bar, int64_t e0, int lane, uint32_t tx) {
    // sparse rows first: they are the slow requests (one 64-byte granule each)
    const int64_t e = e0 + lane;
    const char* p = reinterpret_cast<const char*>(a.rigid_body + e * a.rb_stride + a.rb_off);
    const char* w16 = p - (reinterpret_cast<uintptr_t>(p) & 8u);            // 16-byte aligned window start
    unsigned char* d = sb + PK_IMU + lane * PK_IMU_ROW;
    cp_async16(d, w16); cp_async16(d + 16, w16 + 16); cp_async16(d + 32, w16 + 32);
    const char* pl = reinterpret_cast<const char*>(a.net_contact + e * a.cf_stride + a.cf_l_off);
    const char* pr = reinterpret_cast<const char*>(a.net_contact + e * a.cf_stride + a.cf_r_off);
    const uint32_t ol = (uint32_t)(reinterpret_cast<uintptr_t>(pl) & 8u), orr = (uint32_t)(reinterpret_cast<uintptr_t>(pr) & 8u);
    unsigned char* f = sb + PK_FEET + lane * PK_FEET_ROW;
    cp_async16(f, pl - ol);
    if (ol) cp_async16(f + 16, pl - ol + 16);                                // 12 bytes at +8 of the window spill into the next chunk
    cp_async16(f + 32, pr - orr);
    if (orr) cp_async16(f + 48, pr - orr + 16);
    cp_async_arrive_noinc(bar);
    if (lane == 0) {
        mbar_arrive_expect_tx(bar, tx);
        bulk_g2s(sb + PK_DOF, a.dof_state + e0 * DOF_ROW, PK_WT * 144, bar);
        bulk_g2s(sb + PK_ROOT, a.root_states + e0 * 26, PK_WT * 104, bar);
        bulk_g2s(sb + PK_GOAL, a.goal + e0 * 2, PK_WT * 8, bar);
        bulk_g2s(sb + PK_BINIT, a.ball_init + e0 * 2, PK_WT * 8, bar);
        bulk_g2s(sb + PK_RESET, a.reset_in + e0, PK_WT * 8, bar);
        bulk_g2s(sb + PK_PROG, a.progress_in + e0, PK_WT * 8, bar);
        if (a.prev_lin_vel) bulk_g2s(sb + PK_PREV, a.prev_lin_vel + e0 * 3, PK_WT * 12, bar);
        if (a.values) bulk_g2s(sb + PK_VALUE, a.values + e0, PK_WT * 4, bar);
    }
}

__global__ void __launch_bounds__(PK_THREADS, 1) task_persist_kernel(const TaskArgs a, const __grid_constant__ BezkTaskCfg cfg) {
    extern __shared__ __align__(128) unsigned char pk_smem[];
    __shared__ __align__(8) uint64_t s_full[PK_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles are dealt to the grid's warps round-robin: global warp g = blockIdx.x * PK_C + warp takes tiles g, g + G, ...
    const int64_t ntiles = a.n / PK_WT;
    const int64_t G = (int64_t)gridDim.x * PK_C;
    const int64_t g = (int64_t)warp * gridDim.x + blockIdx.x;          // consecutive tiles land on different SMs
    unsigned char* sb = pk_smem + (size_t)warp * PK_WARP_BYTES;
    float* s_obs = reinterpret_cast<float*>(sb + PK_STAGE_BYTES);
    float* s_dof = reinterpret_cast<float*>(sb + PK_DOF);
    float* s_root = reinterpret_cast<float*>(sb + PK_ROOT);
    uint64_t* bar = &s_full[warp];
    const uint32_t tx = PK_WT * (144 + 104 + 8 + 8 + 8 + 8) + (a.prev_lin_vel ? PK_WT * 12 : 0) + (a.values ? PK_WT * 4 : 0);

    pdl_launch_dependents();
    if (lane == 0) { mbar_init(bar, 1 + 32); fence_mbar_init(); }
    __syncwarp();
    pdl_wait();                                             // nothing above touches global memory
    if (g < ntiles) pk_issue(a, sb, bar, g * PK_WT, lane, tx);

    bool store_pending = false;
    uint32_t phase = 0;
    for (int64_t tile = g; tile < ntiles; tile += G) {
        const int64_t e0 = tile * PK_WT;
        const int64_t e = e0 + lane;
        mbar_wait(bar, phase);
        phase ^= 1u;

        // ---- bookkeeping + masked reset (vec_task.py:331-332, kick_env.py:429-435, 779-850), as in task_tile_kernel ----
        const int64_t reset_prev = reinterpret_cast<const long long*>(sb + PK_RESET)[lane];
        int64_t progress = reinterpret_cast<const long long*>(sb + PK_PROG)[lane];
        int64_t timeout = 0, reset_cur = reset_prev;
        unsigned pending = __ballot_sync(0xffffffffu, reset_prev != 0);
        while (pending) {
            const int r = __ffs(pending) - 1;
            pending &= pending - 1;
            const int64_t env = e0 + r;
            if (lane < 9) {
                float u4[4];
                if (a.uniforms) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) u4[q] = a.uniforms[env * 36 + 4 * lane + q];
                } else {
                    const int64_t genv = a.env_base + env;
                    const Philox4 x = philox4x32_10((uint32_t)genv, (uint32_t)((uint64_t)genv >> 32), (uint32_t)a.step,
                                                    ((uint32_t)(a.step >> 32) << 4) + (uint32_t)lane, (uint32_t)a.seed,
                                                    (uint32_t)(a.seed >> 32));
                    u4[0] = u01(x.x); u4[1] = u01(x.y); u4[2] = u01(x.z); u4[3] = u01(x.w);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int dd = 4 * lane + q;
                    if (dd < 18) {
                        const float off = cfg.reset_pos_span * u4[q] + cfg.reset_pos_lo;
                        s_dof[r * DOF_ROW + 2 * dd] = tensor_clamp(cfg.default_dof_pos[dd] + off, cfg.dof_lower[dd], cfg.dof_upper[dd]);
                    } else {
                        s_dof[r * DOF_ROW + 2 * (dd - 18) + 1] = cfg.reset_vel_span * u4[q] + cfg.reset_vel_lo;
                    }
                }
            }
            __syncwarp();
            if (lane < 9)
                reinterpret_cast<float4*>(a.dof_state_wb + env * DOF_ROW)[lane] = reinterpret_cast<const float4*>(s_dof + r * DOF_ROW)[lane];
            if ((cfg.flags & BEZK_F_RESET_ROOT_STATES) && lane < 26) {
                const float t = a.initial_root[env * 26 + lane];
                s_root[r * 26 + lane] = t;
                a.root_states_wb[env * 26 + lane] = t;
            }
            __syncwarp();
        }
        timeout = (progress >= (int64_t)cfg.max_episode_length - 1) ? 1 : 0;
        progress += 1;
        if (a.randomize_buf) a.randomize_buf[e] += 1;
        if (reset_prev != 0) { progress = 0; reset_cur = 0; }

        // ---- this env's inputs: shared -> registers ----
        float row[36];
        {
            const float4* r4 = reinterpret_cast<const float4*>(s_dof + lane * DOF_ROW);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const float4 t = r4[q];
                row[4 * q] = t.x; row[4 * q + 1] = t.y; row[4 * q + 2] = t.z; row[4 * q + 3] = t.w;
            }
        }
        const float* rr = s_root + lane * 26;
        const float bez[3] = {rr[0], rr[1], rr[2]};
        const float ball_xy[2] = {rr[13], rr[14]}, ball_vxy[2] = {rr[20], rr[21]};
        const float2 g2 = reinterpret_cast<const float2*>(sb + PK_GOAL)[lane];
        const float2 b2 = reinterpret_cast<const float2*>(sb + PK_BINIT)[lane];
        const float goal[2] = {g2.x, g2.y}, binit[2] = {b2.x, b2.y};
        float pv[3];
        const char* psrc = reinterpret_cast<const char*>(a.rigid_body + e * a.rb_stride + a.rb_off);
        const float2* im = reinterpret_cast<const float2*>(sb + PK_IMU + lane * PK_IMU_ROW + ((reinterpret_cast<uintptr_t>(psrc) & 8u) ? 8 : 0));
        const float2 i0 = im[0], i1 = im[1], i2 = im[2], i3 = im[3], i4 = im[4];
        const float q[4] = {i0.x, i0.y, i1.x, i1.y};
        const float v[3] = {i2.x, i2.y, i3.x};
        const float w[3] = {i3.y, i4.x, i4.y};
        if (a.prev_lin_vel) {
            const float* sp = reinterpret_cast<const float*>(sb + PK_PREV) + lane * 3;
            pv[0] = sp[0]; pv[1] = sp[1]; pv[2] = sp[2];
        } else { pv[0] = v[0]; pv[1] = v[1]; pv[2] = v[2]; }                // aliasing, kick_env.py:930
        const char* pls = reinterpret_cast<const char*>(a.net_contact + e * a.cf_stride + a.cf_l_off);
        const char* prs = reinterpret_cast<const char*>(a.net_contact + e * a.cf_stride + a.cf_r_off);
        const float* fls = reinterpret_cast<const float*>(sb + PK_FEET + lane * PK_FEET_ROW + (reinterpret_cast<uintptr_t>(pls) & 8u));
        const float* frs = reinterpret_cast<const float*>(sb + PK_FEET + lane * PK_FEET_ROW + 32 + (reinterpret_cast<uintptr_t>(prs) & 8u));
        const float2 fl2 = *reinterpret_cast<const float2*>(fls), fr2 = *reinterpret_cast<const float2*>(frs);
        const float4 fl4 = make_float4(fl2.x, fl2.y, fls[2], 0.0f), fr4 = make_float4(fr2.x, fr2.y, frs[2], 0.0f);
        const float value = a.values ? reinterpret_cast<const float*>(sb + PK_VALUE)[lane] : 0.0f;
        // ---- everything this warp needs is in registers: refill the stage with the NEXT tile while this one is computed ----
        fence_proxy_async_smem();                           // our generic-proxy reads / patches are ordered before the async-proxy writes
        __syncwarp();
        if (tile + G < ntiles) pk_issue(a, sb, bar, (tile + G) * PK_WT, lane, tx);

        // ---- observations (kick_env.py:749-777) ----
        float imu6[6], orn2[2], feet[8];
        {
            Mth<true> mo;
            imu_term(q, v, w, pv, cfg, imu6, mo);
            off_orn_term(bez[0], bez[1], q, goal[0], goal[1], orn2, mo);
            if (mo.bad()) {
                Mth<false> mp;
                imu_term(q, v, w, pv, cfg, imu6, mp);
                off_orn_term(bez[0], bez[1], q, goal[0], goal[1], orn2, mp);
            }
        }
        if (a.prev_lin_vel) {
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) a.prev_lin_vel[e * 3 + c3] = v[c3];
        }
        {
            const float fl[3] = {fl4.x, fl4.y, fl4.z}, fr[3] = {fr4.x, fr4.y, fr4.z};
            float l3[3] = {fl[0], fl[1], fl[2]}, r3[3] = {fr[0], fr[1], fr[2]};
            float lb[4], rbits[4];
            foot_bits(l3, lb);
            foot_bits(r3, rbits);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) { feet[c4] = lb[c4]; feet[4 + c4] = rbits[c4]; }
            if (cfg.flags & BEZK_F_WRITE_CONTACT_FILTER) {                 // in-place filter, :987-990
                float* cf_l = a.net_contact + e * a.cf_stride + a.cf_l_off;
                float* cf_r = a.net_contact + e * a.cf_stride + a.cf_r_off;
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3) {
                    if (__float_as_uint(l3[c3]) != __float_as_uint(fl[c3])) cf_l[c3] = l3[c3];
                    if (__float_as_uint(r3[c3]) != __float_as_uint(fr[c3])) cf_r[c3] = r3[c3];
                }
            }
        }
        float pos_sq = 0.0f;
#pragma unroll
        for (int j = 0; j < 18; ++j) {
            const float dd = cfg.default_dof_pos[j] - row[2 * j];
            pos_sq += dd * dd;
        }
        // ---- observation rows -> private output buffer -> one bulk store ----
        if (store_pending) {
            if (lane == 0) bulk_wait_read0();              // the previous tile's store has finished reading the buffer
            __syncwarp();
        }
        {
            float2* o2 = reinterpret_cast<float2*>(s_obs + lane * 54);
#pragma unroll
            for (int c9 = 0; c9 < 9; ++c9) {
                o2[c9] = make_float2(row[4 * c9], row[4 * c9 + 2]);
                o2[9 + c9] = make_float2(row[4 * c9 + 1], row[4 * c9 + 3]);
            }
            o2[18] = make_float2(imu6[0], imu6[1]); o2[19] = make_float2(imu6[2], imu6[3]); o2[20] = make_float2(imu6[4], imu6[5]);
            o2[21] = make_float2(orn2[0], orn2[1]);
            o2[22] = make_float2(feet[0], feet[1]); o2[23] = make_float2(feet[2], feet[3]);
            o2[24] = make_float2(feet[4], feet[5]); o2[25] = make_float2(feet[6], feet[7]);
            o2[26] = make_float2(binit[0], binit[1]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(a.obs + e0 * 54, s_obs, PK_OUT_BYTES);
            bulk_commit();
        }
        store_pending = true;

        // ---- reward / termination (overlaps the bulk store) ----
        RewardIn s;
        s.bez[0] = bez[0]; s.bez[1] = bez[1]; s.bez[2] = bez[2];
        s.ball_xy[0] = ball_xy[0]; s.ball_xy[1] = ball_xy[1];
        s.ball_vxy[0] = ball_vxy[0]; s.ball_vxy[1] = ball_vxy[1];
        s.goal[0] = goal[0]; s.goal[1] = goal[1];
        s.ball_init[0] = binit[0]; s.ball_init[1] = binit[1];
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) { s.v[c3] = v[c3]; s.w[c3] = w[c3]; }
        s.pos_sq = pos_sq;
        float rew;
        int64_t reset;
        Mth<true> mr;
        reward_term(s, cfg, progress, reset_cur, &rew, &reset, mr);
        if (mr.bad()) {
            Mth<false> mp;
            reward_term(s, cfg, progress, reset_cur, &rew, &reset, mp);
        }
        a.timeout_buf[e] = timeout;
        a.rew[e] = rew;
        a.reset_out[e] = reset;
        a.progress_out[e] = progress;
        rollout_epilogue(a, e, rew, reset, timeout, value);
    }
    if (store_pending && lane == 0) bulk_wait_read0();      // shared memory must outlive the last bulk store's reads
}

bool persist_eligible(int task, int parts, const TaskArgs& a, const BezkTaskCfg& cfg) {
    // OPT-IN (BEZK_PERSIST=1): measured on B200 (profiles/r02_persist.md) the one-shot tile kernel is as fast or faster at every
    // size -- the step is bound by the per-env dependent instruction chain times the warps an SM can hold (16 there, <= 11 here
    // because of shared memory), not by the latency of its loads -- so the one-shot kernel ships and this one stays as the A/B.
    static const int mode = env_int("BEZK_PERSIST", 0);
    if (!mode) return false;
    if (task != BEZK_TASK_KICK || parts != (BEZK_PART_BOOKKEEP | BEZK_PART_OBS | BEZK_PART_REWARD)) return false;
    if ((cfg.flags & BEZK_F_CLEATS) || a.obs_clipped || !a.use_tma || !a.rb_vec2 || !a.cf_vec2) return false;
    // the 16-byte windows around the sparse rows reach 8 bytes before / after them: Isaac Gym's layout only, and not the last body
    if (a.rb_stride != cfg.num_bodies * 13 || a.cf_stride != cfg.num_bodies * 3) return false;
    if (cfg.imu_body + 1 >= cfg.num_bodies || cfg.left_foot_body + 1 >= cfg.num_bodies || cfg.right_foot_body + 1 >= cfg.num_bodies ||
        cfg.left_foot_body < 1 || cfg.right_foot_body < 1) return false;
    if ((reinterpret_cast<uintptr_t>(a.rigid_body) & 15u) || (reinterpret_cast<uintptr_t>(a.net_contact) & 15u)) return false;
    if (a.n < PK_WT || a.n % PK_WT != 0) return false;
    const auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!al16(a.goal) || !al16(a.ball_init) || !al16(a.reset_in) || !al16(a.progress_in) || !al16(a.prev_lin_vel) || !al16(a.values))
        return false;
    if (a.n < (int64_t)env_int("BEZK_PERSIST_MIN_ENVS", 32)) return false;
    return true;
}

cudaError_t launch_task_persist(const TaskArgs& a, const BezkTaskCfg& cfg, cudaStream_t st) {
    static SmemOptIn opt_in;
    if (cudaError_t err = opt_in.ensure(task_persist_kernel, PK_SMEM)) return err;
    const int64_t ntiles = a.n / PK_WT;
    const unsigned grid = (unsigned)(ntiles < 148 ? ntiles : 148);      // few tiles: one per SM (warp 0 of each CTA) rather than 10 per SM
    return launch_ex(task_persist_kernel, dim3(grid), dim3(PK_THREADS), PK_SMEM, st, a, cfg);
}

}  // namespace bezk
