// Multi-GPU exchange over NVLink peer memory for the k-grid sharded sweep.
//
// One process per GPU.  Every rank cudaMalloc's one arena, exports it with CUDA IPC and maps the
// arenas of its peers, so kernels address peer HBM directly (loads / stores travel over NVLink /
// NVSwitch).  Three device-side primitives replace the NCCL collectives of round 1:
//
//   k_peer_signal   release-store of a monotonically increasing epoch into a flag slot of every
//                   target rank (after a system-scope fence: everything the stream did before,
//                   remote stores included, is visible to whoever acquires the flag)
//   k_peer_wait     ONE thread per source spins (acquire loads, bounded by a time-out) on the local
//                   flag slots; the kernels queued behind it on the stream start once every source
//                   has signalled.  A waiter holds one CTA slot, never the SMs the sweep needs.
//   k_key_merge     the max-with-index reduction of the packed arg-max keys as an in-place
//                   reduce-scatter + all-gather in ONE kernel: rank r owns slice r of the pixels,
//                   reads that slice from every rank (7 remote loads in flight per thread), takes
//                   the 64-bit max and stores the result into every rank's key buffer.  Per rank
//                   2 (W-1)/W of the key bytes cross NVLink, the same volume as a ring all-reduce,
//                   in two NVLink latencies instead of 2 (W-1) ring steps.
//
// The finalize kernel then writes the winner payload of the pixels a rank owns straight into the
// destination rank's output arrays (lockin_finalize.cuh, owner-writes), so no payload reduction exists.
#include "common.cuh"

namespace gpa {

struct PeerPtrs {
    void* p[GPA_MAX_PEERS];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_peer_v2(const ulonglong2* p) {
    ulonglong2 v;   // relaxed system-scope load: never served from a stale L1 line of an earlier step
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// thread t < n_targets: targets.p[t] is the address of MY slot in target t's flag array
__global__ void k_peer_signal(const PeerPtrs targets, int n_targets, unsigned long long epoch) {
    if ((int)threadIdx.x < n_targets) {
        __threadfence_system();
        st_release_sys(static_cast<unsigned long long*>(targets.p[threadIdx.x]), epoch);
    }
}

// thread t < n_sources waits for flags[t] >= epoch; on time-out *status is set and the kernel returns
// (the host raises at its next check; a hung peer must not hang this GPU)
__global__ void k_peer_wait(const unsigned long long* flags, int n_sources, unsigned long long epoch,
                            unsigned long long timeout_ns, int* status) {
    if ((int)threadIdx.x < n_sources) {
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
            if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > timeout_ns) {
                atomicExch(status, 1 + (int)threadIdx.x);
                break;
            }
            __nanosleep(64);
        }
    }
}

// In-place all-reduce (max) of packed keys: this rank reduces elements [lo, hi) of every rank's buffer
template <typename V>
__device__ __forceinline__ V key_max(V a, V b);
template <>
__device__ __forceinline__ ulonglong2 key_max(ulonglong2 a, ulonglong2 b) {
    return make_ulonglong2(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y);
}
template <>
__device__ __forceinline__ unsigned long long key_max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ ulonglong2 ld_any(const ulonglong2* p) { return ld_peer_v2(p); }
__device__ __forceinline__ unsigned long long ld_any(const unsigned long long* p) { return ld_peer(p); }

template <typename V>
__global__ void __launch_bounds__(256) k_key_merge(const PeerPtrs keys, int world, size_t lo, size_t hi) {
    for (size_t i = lo + (size_t)blockIdx.x * 256 + threadIdx.x; i < hi; i += (size_t)gridDim.x * 256) {
        V v[GPA_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < GPA_MAX_PEERS; ++r)
            if (r < world) v[r] = ld_any(static_cast<const V*>(keys.p[r]) + i);
        V m = v[0];
#pragma unroll
        for (int r = 1; r < GPA_MAX_PEERS; ++r)
            if (r < world) m = key_max(m, v[r]);
#pragma unroll
        for (int r = 0; r < GPA_MAX_PEERS; ++r)
            if (r < world) static_cast<V*>(keys.p[r])[i] = m;
    }
}

// w = (wx[row], wy[plane]) of the winner packed in the key (gpa_sweep_argmax), comp_stride elements apart
template <typename R>
__global__ void k_key_to_w(const unsigned long long* __restrict__ key, size_t n, size_t comp_stride,
                           const double* __restrict__ wx, const double* __restrict__ wy, int n_planes, int list_mode,
                           R* __restrict__ w) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = key[i];
        R a = 0, b = 0;
        if ((k >> 32) != 0ull) {
            const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
            const unsigned plane = list_mode ? idx : idx % (unsigned)n_planes;
            const unsigned row = list_mode ? idx : idx / (unsigned)n_planes;
            a = (R)wx[row];
            b = (R)wy[plane];
        }
        w[i] = a;
        w[comp_stride + i] = b;
    }
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle) {
    GPA_REQUIRE(dev_ptr && handle && bytes > 0, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GPA_PEER_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    GPA_CHECK_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("gpa_peer_alloc: %s", cudaGetErrorString(e));
        return GPA_ERR_CUDA;
    }
    std::memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return GPA_OK;
}

extern "C" int gpa_peer_open(const unsigned char* handle, void** dev_ptr) {
    GPA_REQUIRE(dev_ptr && handle, "bad argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    GPA_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GPA_OK;
}

extern "C" int gpa_peer_close(void* dev_ptr) {
    GPA_REQUIRE(dev_ptr, "bad argument");
    GPA_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return GPA_OK;
}

extern "C" int gpa_peer_free(void* dev_ptr) {
    GPA_REQUIRE(dev_ptr, "bad argument");
    GPA_CHECK_CUDA(cudaFree(dev_ptr));
    return GPA_OK;
}

extern "C" int gpa_peer_signal(void* const* target_slots, int n_targets, unsigned long long epoch, void* stream) {
    GPA_REQUIRE(target_slots && n_targets >= 0 && n_targets <= GPA_MAX_PEERS, "bad argument");
    if (n_targets == 0) return GPA_OK;
    PeerPtrs t;
    for (int i = 0; i < n_targets; ++i) {
        GPA_REQUIRE(target_slots[i] != nullptr, "null flag slot");
        t.p[i] = target_slots[i];
    }
    k_peer_signal<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(t, n_targets, epoch);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_peer_wait(const unsigned long long* flags, int n_sources, unsigned long long epoch,
                             double timeout_s, int* status, void* stream) {
    GPA_REQUIRE(flags && status && n_sources >= 0 && n_sources <= 32 && timeout_s > 0, "bad argument");
    if (n_sources == 0) return GPA_OK;
    k_peer_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, n_sources, epoch,
                                                                (unsigned long long)(timeout_s * 1e9), status);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    GPA_REQUIRE(dst && src, "null pointer argument");
    if (bytes == 0) return GPA_OK;
    GPA_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return GPA_OK;
}

extern "C" int gpa_host_register(void* host_ptr, size_t bytes) {
    GPA_REQUIRE(host_ptr && bytes > 0, "bad argument");
    GPA_CHECK_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return GPA_OK;
}

extern "C" int gpa_host_unregister(void* host_ptr) {
    GPA_REQUIRE(host_ptr, "bad argument");
    GPA_CHECK_CUDA(cudaHostUnregister(host_ptr));
    return GPA_OK;
}

extern "C" int gpa_key_merge(void* const* key_ptrs, int world, int rank, size_t n_keys, void* stream) {
    GPA_REQUIRE(key_ptrs && world >= 1 && world <= GPA_MAX_PEERS && rank >= 0 && rank < world, "bad argument");
    if (n_keys == 0 || world == 1) return GPA_OK;
    PeerPtrs k;
    bool vec = n_keys % 2 == 0;
    for (int r = 0; r < world; ++r) {
        GPA_REQUIRE(key_ptrs[r] != nullptr, "null key buffer");
        k.p[r] = key_ptrs[r];
        vec = vec && (reinterpret_cast<uintptr_t>(key_ptrs[r]) % 16 == 0);
    }
    const size_t n = vec ? n_keys / 2 : n_keys;
    const size_t lo = n * (size_t)rank / world, hi = n * (size_t)(rank + 1) / world;
    if (hi == lo) return GPA_OK;
    size_t blocks = (hi - lo + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KernelTimer timer("k_key_merge", st);
    if (vec) k_key_merge<ulonglong2><<<(unsigned)blocks, 256, 0, st>>>(k, world, lo, hi);
    else k_key_merge<unsigned long long><<<(unsigned)blocks, 256, 0, st>>>(k, world, lo, hi);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_key_to_w(const unsigned long long* key, size_t n, size_t comp_stride, const double* wx_dev,
                            const double* wy_dev, int n_planes, int list_mode, int out_f64, void* w, void* stream) {
    GPA_REQUIRE(key && wx_dev && wy_dev && w && n_planes >= 1, "bad argument");
    if (n == 0) return GPA_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_f64) k_key_to_w<double><<<(unsigned)blocks, 256, 0, st>>>(key, n, comp_stride, wx_dev, wy_dev, n_planes, list_mode, static_cast<double*>(w));
    else k_key_to_w<float><<<(unsigned)blocks, 256, 0, st>>>(key, n, comp_stride, wx_dev, wy_dev, n_planes, list_mode, static_cast<float*>(w));
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}
