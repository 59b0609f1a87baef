// Strict binary64 instantiation.  MUST be compiled with -fmad=false (see Makefile): SPEC-PROVISIONAL
// forbids fused multiply-add, and bit-for-bit agreement with the CPU oracle depends on it.
#include "nt_trace.cuh"

int nt_launch_render_f64(const NtDevScene &s, const NtRenderArgs &a, void *stream) {
    return nt::launch_render<double>(s, a, (cudaStream_t)stream);
}
int nt_launch_trace_f64(const NtDevScene &s, const NtTraceArgs &a, void *stream) {
    return nt::launch_trace<double>(s, a, (cudaStream_t)stream);
}
