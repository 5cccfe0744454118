// Host-side helpers: error reporting, TMA tensor-map encoding (driver entry point fetched at run time,
// so the library does not link against libcuda), SM count.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <set>
#include <string>
#include <utility>

namespace tvae {

inline thread_local std::string g_last_error;
inline int g_dev_knob[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // development knobs (tvae_test_set_knob): 0 = conv1_wgrad reduction splits
inline std::atomic<long long> g_launch_count{0};   // kernels launched by this library (bench.py reports it)

inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define TVAE_CHECK_CUDA(expr)                                                                          \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return ::tvae::fail(-2, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

#define TVAE_REQUIRE(cond, msg)                                                                        \
    do {                                                                                               \
        if (!(cond)) return ::tvae::fail(-1, std::string("invalid argument: ") + (msg));               \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D 16-bit (fp16 / bf16 bit patterns) row-major tensor [rows][cols], K-major operand boxes {64 cols, box_rows}, 128 B swizzle.
inline int make_tmap_2d_h(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return fail(-1, "TMA operand must be 16-byte aligned with 16-byte row pitch");
    if (box_rows == 0 || box_rows > 256) return fail(-1, "TMA box rows out of range");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (16-bit) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// 2-D 16-bit row-major tensor [rows][cols] read as an MN-major operand: boxes {32 cols (64 B), box_rows}, 64 B swizzle.
inline int make_tmap_2d_mn_h(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return fail(-1, "TMA operand must be 16-byte aligned with 16-byte row pitch");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (16-bit, MN-major) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// MN-major operand view of a row-major 16-bit matrix [rows][cols] (cols % 32 == 0): 3-D map {32 cols, rows, cols/32 column
// blocks}, boxes {32, box_rows, nb}: one TMA fills nb consecutive [box_rows][64 B] column blocks in the 64 B swizzle.
inline int make_tmap_3d_mn_h(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t nb) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15) || (cols & 31)) return fail(-1, "TMA operand must be 16-byte aligned with whole 32-column blocks");
    cuuint64_t dims[3] = {32, rows, cols / 32};
    cuuint64_t strides[2] = {ld * 2, 32 * 2};
    cuuint32_t box[3] = {32, box_rows, nb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (3-D, 16-bit) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// Same view in 64-column blocks (cols % 64 == 0): 3-D map {64 cols, rows, cols/64}, boxes {64, box_rows, nb}: nb consecutive
// [box_rows][128 B] column blocks in the 128 B swizzle (the MN-major layout the tensor core fetches with full 128-byte rows).
inline int make_tmap_3d_mn128_h(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t nb) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15) || (cols & 63)) return fail(-1, "TMA operand must be 16-byte aligned with whole 64-column blocks");
    cuuint64_t dims[3] = {64, rows, cols / 64};
    cuuint64_t strides[2] = {ld * 2, 64 * 2};
    cuuint32_t box[3] = {64, box_rows, nb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (3-D, 16-bit, 128 B blocks) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// Store view of a 16-bit activation tensor [outer][rows][cols] (cols contiguous): boxes {64 cols (128 B), box_rows, 1},
// 128 B swizzle; rows >= `rows` of a box are clipped by the TMA unit (ragged last tile of an image).
inline int make_tmap_3d_store_h(CUtensorMap* tm, void* ptr, uint64_t outer, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols & 63)) return fail(-1, "TMA store target must be 16-byte aligned with whole 64-column blocks");
    cuuint64_t dims[3] = {cols, rows, outer};
    cuuint64_t strides[2] = {cols * 2, rows * cols * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (3-D store) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// fp32 matrix [rows][cols] (row pitch ld elements) as the target of TMA reduce-adds: boxes {128 cols, 32 rows}, no swizzle;
// the box is clipped at (cols, rows).  Returns 1 (not an error) when the matrix does not meet TMA's alignment rules.
inline int make_tmap_2d_f32_reduce(CUtensorMap* tm, void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(-3, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 4) & 15)) return 1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {128, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (fp32 reduce target) failed with code " + std::to_string(static_cast<int>(r)));
    return 0;
}

// ---- optional per-kernel timing (bench.py): CUDA events recorded around every tc_gemm launch on its own stream
struct KernelTimer {
    static constexpr int kMaxNames = 48;
    static constexpr int kMaxEvents = 16384;
    std::atomic<bool> enabled{false};
    std::mutex mu;              // begin() / collect are called from the caller thread (forward) AND the autograd thread (backward)
    int n_names = 0;
    const char* names[kMaxNames];
    int n_events = 0;
    cudaEvent_t start[kMaxEvents], stop[kMaxEvents];
    int name_of[kMaxEvents];
    int n_created = 0;
    int name_id(const char* nm) {       // caller holds `mu`
        for (int i = 0; i < n_names; ++i)
            if (names[i] == nm || strcmp(names[i], nm) == 0) return i;
        if (n_names >= kMaxNames) return -1;
        names[n_names] = nm;
        return n_names++;
    }
    // returns slot or -1.  The slot (and with it the event pair) is reserved under the lock; recording the events needs
    // no lock: each slot is used by exactly one launch.
    int begin(const char* nm, cudaStream_t st) {
        if (!enabled.load(std::memory_order_relaxed)) return -1;
        int e;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (n_events >= kMaxEvents) return -1;
            const int id = name_id(nm);
            if (id < 0) return -1;
            e = n_events++;
            if (e >= n_created) {
                cudaEventCreate(&start[e]);
                cudaEventCreate(&stop[e]);
                n_created = e + 1;
            }
            name_of[e] = id;
        }
        cudaEventRecord(start[e], st);
        return e;
    }
    void end(int e, cudaStream_t st) {
        if (e >= 0) cudaEventRecord(stop[e], st);
    }
};
inline KernelTimer g_timer;

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
inline int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per (kernel, device): remembered per pair so that
// a process driving several GPUs configures every one of them (the library keeps no other mutable global state).
inline cudaError_t smem_optin(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kernel, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.insert({kernel, dev});
    return e;
}

inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace tvae
