// nt_sync.cuh — frame synchronisation between the GPUs of a sharded render: sequence-numbered flags in
// device memory of the gathering GPU (mapped into the other processes through CUDA IPC) or in pinned host memory.
//   post at start   "everything enqueued on this stream before this frame has completed" (rank 0's acknowledgement that
//                   it has consumed the previous frames: a peer may overwrite the frame buffer of two frames ago)
//   wait            spin until a flag has reached a sequence number (bounded: ~2 s, then the time-out counter is set
//                   and the kernel carries on - a lost peer must not hang the GPU)
//   post when done  release-store after the frame's last pixel store, so that whoever acquires the flag sees the pixels
// These replace the NCCL all-reduce that ordered "all shards written" in round 1 (49 us per frame; a flag store over
// NVLink and a local spin are ~2-3 us).  No reference file to cite (/root/reference/README:1-3 holds no code).
#pragma once
#include <cuda_runtime.h>

namespace nt {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifndef NT_SYNC_TIMEOUT_NS
#define NT_SYNC_TIMEOUT_NS 2000000000ull
#endif
// Returns false on time-out.  Sequence numbers wrap: compare by signed difference.
__device__ __forceinline__ bool spin_until(const unsigned *p, unsigned value) {
    if ((int)(ld_acquire_sys(p) - value) >= 0) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        __nanosleep(100);
        if ((int)(ld_acquire_sys(p) - value) >= 0) return true;
        if (global_ns() - t0 > NT_SYNC_TIMEOUT_NS) return false;
    }
}

// Start of a frame's first kernel; called by every thread, before a block barrier that precedes the first pixel store.
__device__ __forceinline__ void frame_sync_begin(const NtRenderArgs &a) {
    if (threadIdx.x == 0) {
        if (a.sync_post_ptr && blockIdx.x == 0) st_release_sys(a.sync_post_ptr, a.sync_post_val);
        if (a.sync_wait_ptr && !spin_until(a.sync_wait_ptr, a.sync_wait_val))
            atomicAdd(a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS + 2, 1ull);
    }
}
// End of a frame's last kernel; called by every thread after its last pixel store.  The last block to arrive publishes
// the flag: every thread fences its own stores at system scope, the block barrier and the device-scope atomic chain
// them to the publishing thread (fences are cumulative).
__device__ __forceinline__ void frame_sync_end(const NtRenderArgs &a) {
    if (!a.sync_done_ptr) return; // kernel parameter: uniform
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long arrived = atomicAdd(a.counters + NT_COUNTER_SLOTS * NT_NCOUNTERS + 1, 1ull);
        if (arrived + 1 == gridDim.x) {
            __threadfence_system();
            st_release_sys(a.sync_done_ptr, a.sync_done_val);
        }
    }
}

// BVH scenes render a frame with many kernels: the flags are handled by this one-thread kernel before (post, wait)
// and after (done) the pipeline.  Also the body of nt_flags_wait_device: lane i spins on flag i.
static __global__ void sync_kernel(unsigned *post_ptr, unsigned post_val, const unsigned *wait_ptr, unsigned wait_n, unsigned wait_val,
                            unsigned *done_ptr, unsigned done_val, unsigned long long *timeouts) {
    if (post_ptr && threadIdx.x == 0) st_release_sys(post_ptr, post_val);
    if (wait_ptr)
        for (unsigned i = threadIdx.x; i < wait_n; i += blockDim.x)
            if (!spin_until(wait_ptr + i, wait_val) && timeouts) atomicAdd(timeouts, 1ull);
    __syncthreads();
    if (done_ptr && threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(done_ptr, done_val);
    }
}

} // namespace nt
