// Fast binary32 instantiation (SPEC-PROVISIONAL §7): FMA contraction and -use_fast_math.
#include "nt_trace.cuh"

int nt_launch_render_f32(const NtDevScene &s, const NtRenderArgs &a, void *stream) {
    return nt::launch_render<float>(s, a, (cudaStream_t)stream);
}
int nt_launch_trace_f32(const NtDevScene &s, const NtTraceArgs &a, void *stream) {
    return nt::launch_trace<float>(s, a, (cudaStream_t)stream);
}
size_t nt_flat_smem_bytes(const NtDevScene &s, int precision) {
    return precision == 0 ? nt::flat_smem_bytes<double>(s, s.use_bvh != 0)
                          : nt::flat_smem_bytes<float>(s, s.use_bvh != 0);
}
size_t nt_sample_buffer_bytes(const NtDevScene &s, const NtRenderArgs &a, int precision) {
    if (!s.use_bvh) return 0;
    const size_t n = (size_t)a.tiles_x * a.tiles_y * (a.spp / a.lanes) * 32;
    return n * 3 * (precision == 0 ? sizeof(double) : sizeof(float));
}
size_t nt_wavefront_min_bytes(const NtRenderArgs &a, int precision) {
    const size_t per = precision == 0 ? nt::wf_bytes_per_sample<double>(a.max_depth) : nt::wf_bytes_per_sample<float>(a.max_depth);
    return 512 + 256 * 8 * (size_t)a.max_depth + 32 * per + 4096; // header (nt_wavefront.cuh launch_wavefront) + alignment + one warp of samples
}
size_t nt_wavefront_bytes(const NtDevScene &s, const NtRenderArgs &a, int precision) {
    return precision == 0 ? nt::wavefront_bytes<double>(s, a) : nt::wavefront_bytes<float>(s, a);
}
