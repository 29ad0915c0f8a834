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
// cudaGetLastError() after a <<<>>> launch -> 0 or a recorded error
int check_launch(const char* what);

int sm_count();

// TMA descriptor of a 16-bit row-major matrix [outer][inner] (leading dimension `ld` elements), box {box_inner, box_outer}
int make_tmap_16bit(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                    int box_outer, CUtensorMapSwizzle swizzle);

// Optional per-launch timing of the GEMM kernel with CUDA events on the launching stream (bench.py's roofline).
bool gemm_timing_enabled();
void gemm_timing_record(cudaStream_t st, bool begin);

}  // namespace p2t
