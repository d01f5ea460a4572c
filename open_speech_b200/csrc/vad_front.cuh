// Fused Silero-VAD front end (csrc/vad_front.cu): one persistent tcgen05 kernel from samples to LSTM gate pre-activations.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace osb {

struct VadFront;  // device-resident weight image in the kernel's consumption order

// float offsets of the tensors inside the flat host weight blob (vad/silero.py WEIGHT_LAYOUT)
struct VadFrontLayout {
    size_t basis, e1w, e1b, e2w, e2b, e3w, e3b, e4w, e4b, wih, bih, bhh;
};

int vad_front_create(const float* weights_host, const VadFrontLayout& layout, VadFront** out);
void vad_front_destroy(VadFront* f);
// windows are rows (stream, t): window w of the launch = stream w / wins_per_stream, window win0 + w % wins_per_stream of that stream;
// d_pre (ceil(total_windows / 128) * 65536 floats) receives W_ih.x + b_ih + b_hh in the interleaved layout of vad.cu's pre_at().
// tile_begin / tile_end (-1: to the last tile) restrict the launch to a range of the chunk's 128-window tiles; max_ctas (0: one per SM)
// caps the grid, so that the launch fits on the SMs a concurrently running recurrence leaves free (vad.cu, pipelined chunks)
int vad_front_tiles(long long total_windows);
int launch_vad_front_fused(const VadFront* f, const void* d_audio, int fmt, long long audio_stride, int wins_per_stream, long long win0,
                           long long total_windows, float* d_pre, cudaStream_t st, int tile_begin = 0, int tile_end = -1, int max_ctas = 0);

}  // namespace osb
