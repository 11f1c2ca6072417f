// step2d_acc_pipe.cuh -- ALL time steps of a single-GPU 2D FAST run in ONE launch of the face-flux kernel (step2d_acc.cuh).
//
// One launch per step leaves three things on the table that no amount of tuning inside the step removes: the ramp-up of a
// fresh grid (every warp starts with an empty TMA ring and waits a full HBM latency), the tail (the last wave of blocks
// runs on a half-empty GPU) and the launch gap.  At 4096^2 that is 6-8 % of a 92 us step, at 256^2 it is most of it.
// Here the grid is launched once with as many one-warp blocks as are resident at a time, and the blocks draw work items
// (step, chunk, tile) IN ORDER from one atomic counter -- all items of step s before any item of step s+1.  An item of step
// s+1 may start as soon as the chunk rows c-1, c, c+1 of step s are complete (its x-halo rows and, because the state
// ping-pongs between two buffers, exactly the items that still read the rows it is going to overwrite):
//
//   finish item (s, c, t):  every lane __threadfence(); lane 0  atomicAdd(done[c], 1)        (done[] only ever grows)
//   start  item (s, c, t):  lane 0 spins on  ld.acquire.gpu done[c'] >= s * ntiles,  c' = c-1, c, c+1
//
// so step s+1 trails step s by the ~35 chunk rows that are in flight and the GPU never drains between steps.  Deadlock
// freedom: an item only waits for items with smaller indices, and those were drawn earlier, i.e. by blocks that are
// running; the TMA producer lane, which runs `stages` boxes AHEAD of its warp and therefore looks at the next item while
// the current one is unfinished, never blocks on a dependency -- it simply arms nothing until the dependency holds, and
// the warp only spins for a box once it has nothing else left to do.  Every spin has a wall-clock timeout that sets the
// context's error word (SHLL_E_TIMEOUT) and lets the kernel finish.
//
// The arithmetic is acc_row() of step2d_acc.cuh, unchanged: results are bitwise those of the one-launch-per-step kernel
// (tests/test_gpu_parity.py).  Multi-GPU slabs keep one launch per step (their halo flags are per-step kernel parameters).
#pragma once
#include "persist1d.cuh"  // ld_acquire_gpu
#include "step2d_acc.cuh"

namespace shll {

struct Step2DPipeParams {
    Step2DTmaParams even;   // step 0, 2, ...: tensor map of the buffer they READ, out[] = planes they WRITE
    CUtensorMap tmap_odd;   // steps 1, 3, ...: they read what the even steps wrote
    int out_delta;          // floats from an even step's out[k] to an odd step's out[k] (the other ping-pong buffer)
    int nsteps;
    int nitems;             // items per step = ntiles * nchunks
    unsigned *work;         // [0] items drawn so far; zeroed by the host before the launch
    unsigned *done;         // [nchunks] completed items per chunk row, all steps added up; zeroed by the host
    unsigned *err;
    unsigned long long timeout_ns;
};

struct PipeProducer {
    int items[16];          // FIFO of the items this warp consumes, by sequence number & 15
    int valid;              // `item` has boxes left to arm
    int box, nboxes, x0, ybase, odd;
    int dep_ok, dep_lo, dep_hi;
    unsigned dep_target;
    int stage, seq;
    int armed, consumed;    // boxes armed / fully read so far (stream positions)
};

__device__ __forceinline__ void pipe_item_decode(const Step2DPipeParams &T, int item, int &step, int &tile, int &chunk, int &r0, int &r1)
{
    const Step2DParams &P = T.even.base;
    step = item / T.nitems;
    const int rem = item - step * T.nitems;
    chunk = rem / P.ntiles;
    tile = rem - chunk * P.ntiles;
    r0 = (int)(((long)chunk * P.nx) / P.nchunks);
    r1 = (int)(((long)(chunk + 1) * P.nx) / P.nchunks);
}

// lane 0 only
template <int ORDER>
static __device__ __noinline__ void pipe_open(const Step2DPipeParams *T, PipeProducer *Q, int item)
{
    int step, tile, chunk, r0, r1;
    pipe_item_decode(*T, item, step, tile, chunk, r0, r1);
    const int xs = tile * 60 - 2;  // USEFUL = 60 columns, HL * VEC = 2 (step2d_acc.cuh)
    Q->x0 = xs & ~3;
    Q->ybase = r0 - ORDER + 2;
    Q->nboxes = (r1 - r0 + 2 * ORDER - 1) / 4 + 1;
    Q->box = 0;
    Q->odd = step & 1;
    Q->dep_ok = (step == 0);
    Q->dep_lo = max(chunk - 1, 0);
    Q->dep_hi = min(chunk + 1, T->even.base.nchunks - 1);
    Q->dep_target = (unsigned)step * (unsigned)T->even.base.ntiles;
    Q->valid = 1;
}

// lane 0 only: arm the next box if there is one, a ring stage is free and its item's dependencies hold.  Returns whether
// it armed.
template <int ORDER>
static __device__ __noinline__ bool pipe_try_arm(const Step2DPipeParams *T, PipeProducer *Q, uint32_t ring, uint32_t bars, int stages,
                                                 uint32_t stage_stride, uint32_t stage_bytes)
{
    if (!Q->valid || Q->armed >= Q->consumed + stages) return false;
    if (!Q->dep_ok) {
        // three independent relaxed loads (one round trip to L2, not three), then one acquire fence if they all pass
        const int mid = min(Q->dep_lo + 1, Q->dep_hi);
        unsigned v0, v1, v2;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(T->done + Q->dep_lo) : "memory");
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(T->done + mid) : "memory");
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(T->done + Q->dep_hi) : "memory");
        const unsigned t = Q->dep_target;
        if ((int)(v0 - t) < 0 || (int)(v1 - t) < 0 || (int)(v2 - t) < 0) return false;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        Q->dep_ok = 1;
        asm volatile("fence.proxy.async.global;" ::: "memory");  // the rows were written through the generic proxy, TMA reads them
    }
    const int stage = Q->stage;
    const uint32_t bar = bars + 8u * stage;
    mbar_expect_tx(bar, stage_bytes);
    tma_load_3d(ring + stage_stride * stage, Q->odd ? &T->tmap_odd : &T->even.tmap, Q->x0, Q->ybase + Q->box * 4, 0, bar);
    Q->stage = (stage + 1 == stages) ? 0 : stage + 1;
    Q->armed++;
    if (++Q->box == Q->nboxes) {  // that item is fully requested: draw the next one
        const int total = T->nitems * T->nsteps;
        const int next = (int)atomicAdd(T->work, 1u);
        Q->items[(++Q->seq) & 15] = next < total ? next : total;
        if (next < total) pipe_open<ORDER>(T, Q, next);
        else Q->valid = 0;
    }
    return true;
}

// lane 0 only: after a box has been read, refill; before a box is read, make sure it has been requested
template <int ORDER>
__device__ __forceinline__ void pipe_refill(const Step2DPipeParams *T, PipeProducer *Q, uint32_t ring, uint32_t bars, int stages,
                                            uint32_t stage_stride, uint32_t stage_bytes)
{
    Q->consumed++;
    while (pipe_try_arm<ORDER>(T, Q, ring, bars, stages, stage_stride, stage_bytes)) {}
}
template <int ORDER>
static __device__ __noinline__ void pipe_need_box(const Step2DPipeParams *T, PipeProducer *Q, uint32_t ring, uint32_t bars, int stages,
                                                  uint32_t stage_stride, uint32_t stage_bytes)
{
    if (Q->armed > Q->consumed) return;
    const unsigned long long t0 = globaltimer_ns();
    unsigned polls = 0;
    while (!pipe_try_arm<ORDER>(T, Q, ring, bars, stages, stage_stride, stage_bytes)) {
        __nanosleep(64);
        if ((++polls & 255u) == 0 && globaltimer_ns() - t0 > T->timeout_ns) {
            atomicExch(T->err, 1u);
            Q->dep_ok = 1;  // give up on the dependency: the run is reported as failed, but it terminates
        }
    }
}

template <int ORDER, int BC, int LIM, bool WALLTILE, int STASH, class Ctx>
__device__ __forceinline__ void pipe_march(Ctx &X, const AccRows &W, const Step2DPipeParams *T, int rbeg, int rlast, int lane, PipeProducer *Q,
                                           int &stage, uint32_t &parity)
{
    AccState<ORDER> S;
#pragma unroll
    for (int k = 0; k < 4; k++)
        S.fpP[k] = S.D[k] = S.accB[k] = S.gP[k] = S.ep[k] = S.em[k] = S.PhiP[k] = S.acc0[k] = v2bc(0.0f);
    float uin[2][4];
    const int nx = X.P->nx;
    for (int box = 0; box < X.nboxes; box++) {
        const int r = rbeg + 4 * box;
        if (lane == 0) pipe_need_box<ORDER>(T, Q, X.ring, X.bars, X.stages, Ctx::STAGE_STRIDE, Ctx::STAGE_BYTES);
        __syncwarp();
        mbar_wait(X.bars + 8u * stage, parity);
        const int ifirst = r - ORDER;  // rows finished by this box: ifirst .. ifirst + 3
        const bool edge = (r <= 1) || (r + 3 >= nx - 1) || (r + 3 > rlast) || (ifirst < W.r0);
        if (edge) {
#pragma unroll 1
            for (int w = 0; w < 4; w++) {
                if (r + w > rlast) break;
                acc_read_row(X, stage, w, uin);
                acc_row<ORDER, BC, LIM, WALLTILE, true, STASH>(X, W, S, r + w, uin);
            }
        } else {
            X.template read_row<0>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH>(X, W, S, r, uin);
            X.template read_row<1>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH>(X, W, S, r + 1, uin);
            X.template read_row<2>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH>(X, W, S, r + 2, uin);
            X.template read_row<3>(stage, uin);
            acc_row<ORDER, BC, LIM, WALLTILE, false, STASH>(X, W, S, r + 3, uin);
        }
        __syncwarp();  // the box has been fully read
        if (lane == 0) pipe_refill<ORDER>(T, Q, X.ring, X.bars, X.stages, Ctx::STAGE_STRIDE, Ctx::STAGE_BYTES);
        stage++;
        if (stage == X.stages) { stage = 0; parity ^= 1u; }
    }
}

template <int ORDER, int BC, int LIM, int MINB, int STASH>
__global__ void __launch_bounds__(32, MINB) step2d_acc_pipe_kernel(const __grid_constant__ Step2DPipeParams T)
{
    constexpr int R = 4, VEC = 2, HL = 1;
    constexpr int USEFUL = (32 - 2 * HL) * VEC;
    static_assert(USEFUL == 60 && HL * VEC == 2, "pipe_open hard-codes the tile geometry");
    extern __shared__ __align__(128) unsigned char smem[];
    const Step2DParams &P = T.even.base;
    const int lane = threadIdx.x;
    const int nx = P.nx;
    const int never = -(1 << 30);

    TmaCtx<VEC, R> X;
    X.P = &P;
    X.T = &T.even;
    X.ny = P.ny;
    X.wall_lo_row = X.wall_hi_row = X.first_real_row = X.last_real_row = never;  // (window-kernel fields, unused here)
    X.peer_lo_end = 0;                                                            // single GPU: no peer halos
    X.peer_hi_begin = 0x7fffffff;
    X.stages = T.even.stages;
    X.ring = smem_u32(smem);
    X.bars = X.ring + TmaCtx<VEC, R>::STAGE_STRIDE * X.stages;
    AccRows W;
    W.rmin = 0;            // both x ends of the domain are physical walls
    W.rmax = nx - 1;
    W.wall_lo_row = 0;
    W.wall_hi_row = nx - 1;
    W.noslope_lo = 0;
    W.noslope_hi = nx - 1;
    W.quarter = P.quarter;
    W.nquarter = -P.quarter;
    W.stash = 0;
    // shared memory after the ring: mbarriers, [stash], producer state (sized by the host: shll_capi.cu make_tensor_maps)
    uint32_t after = X.bars + 8u * X.stages;
    if (STASH) {
        W.stash = ((after + 15u) & ~15u) + 8u * lane;
        after = ((after + 15u) & ~15u) + (STASH >= 2 ? 4096u : 2048u);
    }
    PipeProducer *Q = reinterpret_cast<PipeProducer *>(smem + (((after + 15u) & ~15u) - X.ring));
    if (lane == 0) {
        for (int s = 0; s < X.stages; s++) mbar_init(X.bars + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        Q->stage = 0;
        Q->seq = 0;
        Q->armed = 0;
        Q->consumed = 0;
        // every item, the first one included, is drawn from the counter: an item is then always held by a block that is
        // running, whatever else shares the GPU (the in-order argument above needs exactly that)
        const int total0 = T.nitems * T.nsteps;
        const int first = (int)atomicAdd(T.work, 1u);
        Q->items[0] = first < total0 ? first : total0;
        Q->valid = 0;
        if (first < total0) pipe_open<ORDER>(&T, Q, first);
        while (pipe_try_arm<ORDER>(&T, Q, X.ring, X.bars, X.stages, TmaCtx<VEC, R>::STAGE_STRIDE, TmaCtx<VEC, R>::STAGE_BYTES)) {}
    }

    const int total = T.nitems * T.nsteps;
    int stage = 0;
    uint32_t parity = 0;
    for (int seq = 0;; seq++) {
        __syncwarp();  // lane 0's FIFO writes are visible
        const int item = *reinterpret_cast<volatile int *>(&Q->items[seq & 15]);
        if (item >= total) break;
        int step, tile, chunk;
        pipe_item_decode(T, item, step, tile, chunk, X.r0, X.r1);
        const int xs = tile * USEFUL - HL * VEC;
        X.x0 = xs & ~3;
        X.j0 = xs + lane * VEC;
        X.owner = (lane >= HL) && (lane < 32 - HL) && (X.j0 < X.ny);
        X.Y.tile_has_wall = (tile == 0) || (tile == P.ntiles - 1);
        X.Y.ghost_lo = (X.j0 + VEC - 1 == -1);
        X.Y.ghost_hi = (X.j0 == X.ny);
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            X.Y.y_inner[v] = (X.j0 + v > 0 && X.j0 + v < X.ny - 1);
            X.Y.outside[v] = (X.j0 + v < 0 || X.j0 + v >= X.ny);
        }
        if (step & 1) X.j0 += T.out_delta;  // from here on j0 only addresses the stores: odd steps write the other buffer
        W.r0 = X.r0;
        W.r1 = X.r1;
        const int rbeg = X.r0 - ORDER;
        const int rlast = X.r1 - 1 + ORDER;
        X.lane_off = (uint32_t)(xs - X.x0 + lane * VEC) * 4u;
        X.ybase = rbeg + 2;
        X.nboxes = (rlast - rbeg) / R + 1;
        if (STASH)
            for (int i = 0; i < (STASH >= 2 ? 16 : 8); i++) stash_store(W.stash + i * 256, v2bc(0.0f));

        if (X.Y.tile_has_wall) pipe_march<ORDER, BC, LIM, true, STASH>(X, W, &T, rbeg, rlast, lane, Q, stage, parity);
        else pipe_march<ORDER, BC, LIM, false, STASH>(X, W, &T, rbeg, rlast, lane, Q, stage, parity);

        // publish: the item's rows are in L2 for every SM before its chunk row's counter moves
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(T.done + chunk, 1u);
    }
}

}  // namespace shll
