// Library-internal helpers: error reporting and launch accounting behind the C ABI (include/p2t_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace p2t {

// Records a message retrievable through p2t_last_error(); returns `code` so callers can `return set_error(...)`.
int set_error(int code, const char* fmt, ...);
// Every kernel launch made by the library bumps this counter (bench.py reports it as gpu_launches).
void count_launch();
// cudaGetLastError() after a <<<>>> launch on `st` -> 0 or a recorded error; counts the launch and, when launch
// timing is on, stamps it (see stamp_launch)
int check_launch(const char* what, cudaStream_t st);
// Per-launch timing of EVERY kernel of the library (bench.py --stages): an event is recorded on the launching stream
// after each launch; consecutive stamps on an in-order stream bracket one kernel (plus its launch gap), and a
// "mark" stamp (p2t_launch_timing_mark) opens each step.  Off by default; not for use under stream capture.
void stamp_launch(const char* name, cudaStream_t st);
// Optional "begin" stamp right before a launch: the kernel is then timed from it (one event, recorded when the
// predecessor has finished) instead of from the previous kernel's end stamp, which would charge it the gap.
inline void stamp_begin(cudaStream_t st) { stamp_launch("begin", st); }

int sm_count();

// TMA descriptor of a 16-bit row-major matrix [outer][inner] (leading dimension `ld` elements), box {box_inner, box_outer}
int make_tmap_16bit(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                    int box_outer, CUtensorMapSwizzle swizzle);

// Optional per-launch timing of the GEMM kernel with CUDA events on the launching stream (bench.py's roofline).
bool gemm_timing_enabled();
void gemm_timing_record(cudaStream_t st, bool begin);

}  // namespace p2t
