// halo_sync.cuh -- flag protocol for the halo exchange that is fused into the step kernels.
//
// Slabs are contiguous along x.  A slab's edge warps store their freshly updated edge
// rows/cells twice: into the local output plane and, through a peer-mapped pointer (CUDA IPC or same-process
// peer access, NVLink 5 / NVSwitch), straight into the neighbour's halo rows.  No NCCL call, no staging copy,
// no extra kernel.  Ordering is by one 32-bit flag per side:
//
//   producer (edge warp, after its stores):  __threadfence_system(); arrive on a local counter;
//       the last edge warp of the step does   st.release.sys  peer_flag = state_index + 1
//   consumer (edge warp, before its loads):   spin on  ld.acquire.sys  own_flag >= state_index + 1
//
// state_index counts time steps since creation (monotonic, never reset), so flags only ever grow.
// WAR safety: a slab signals its neighbour only after *all* its edge warps finished the step, i.e. after they
// are done reading the halo rows the neighbour will overwrite two steps later; the neighbour cannot start that
// later step before it has seen this signal.
//
// A consumer that waits longer than `timeout_ns` records an error and lets the kernel finish (results are then
// garbage, the host reports SHLL_E_TIMEOUT): a lost neighbour must never hang the GPU.
#pragma once
#include <cuda_runtime.h>

namespace shll {

struct HaloSync {
    int enabled;
    unsigned want;             // flag value that means "input halo for this step has arrived"  (= state_index + 1)
    unsigned post;             // flag value to publish when this step's edge output has landed (= state_index + 2)
    unsigned epoch;            // 1-based index of this launch among all step launches of the context
    const unsigned *wait_lo;   // local flags, written remotely by the lower / upper neighbour
    const unsigned *wait_hi;
    unsigned *sig_lo;          // neighbour's flags (peer memory), NULL at a wall
    unsigned *sig_hi;
    unsigned *cnt_lo;          // local arrival counters of edge warps (monotonic)
    unsigned *cnt_hi;
    unsigned edge_warps_lo;    // edge warps per step on each side
    unsigned edge_warps_hi;
    unsigned *err;             // local error word (0 = ok)
    unsigned long long timeout_ns;
    unsigned long long *wait_ns;  // local statistics: [0] ns spent spinning on the lower flag, [1] upper, [2] number of waits that spun
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Whole warp calls this; lane 0 spins.  The error word is sticky: once ANY wait of this context has timed out (a dead
// neighbour never posts a later flag either), every later wait -- the other edge warps of this step and all the steps
// already enqueued behind it -- returns at once, so a lost neighbour costs ONE timeout, not one per edge warp per step.
// The halo rows come from the peer GPU through the generic proxy; the TMA kernels read them through the async proxy, so
// the acquire is followed by a proxy fence (edge warps only: 2*ntiles warps per step).
__device__ __forceinline__ void halo_wait(const HaloSync &S, const unsigned *flag, int side = 0)
{
    if (flag == nullptr) return;
    if ((threadIdx.x & 31) == 0) {
        if ((int)(ld_acquire_sys(flag) - S.want) < 0 && *(volatile unsigned *)S.err == 0u) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned polls = 0;
            while ((int)(ld_acquire_sys(flag) - S.want) < 0) {
                __nanosleep(32);
                if ((++polls & 15u) == 0) {
                    if (*(volatile unsigned *)S.err != 0u) break;
                    if ((polls & 255u) == 0 && globaltimer_ns() - t0 > S.timeout_ns) {
                        atomicExch(S.err, 1u);
                        break;
                    }
                }
            }
            if (S.wait_ns != nullptr) {  // attribution of what is left of the scaling loss (bench.py: halo_wait_us_per_step)
                atomicAdd(S.wait_ns + side, globaltimer_ns() - t0);
                atomicAdd(S.wait_ns + 2, 1ull);
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
}

// Whole warp calls this after its last edge store.
__device__ __forceinline__ void halo_arrive(const HaloSync &S, unsigned *cnt, unsigned edge_warps, unsigned *peer_flag)
{
    if (peer_flag == nullptr) return;
    __threadfence_system();  // every lane: its own peer stores are performed system-wide before the arrival below
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const unsigned prev = atomicAdd(cnt, 1u);
        if (prev + 1u == edge_warps * S.epoch) {
            __threadfence_system();
            st_release_sys(peer_flag, S.post);
        }
    }
}

// Programmatic dependent launch (single-GPU step kernels).  A kernel launched with launch_pdl(..., pdl = true) may be placed
// on the SMs while its predecessor in the stream is still running; it must call this before it touches the state: it lets
// ITS successor be placed as soon as every block of this grid has started, then waits until the predecessor has completed
// and its stores are visible.  Everything before the call may only touch registers and shared memory.
__device__ __forceinline__ void pdl_wait_for_previous_step(int pdl)
{
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
}

}  // namespace shll
