// Probe: which TMA box configurations work for the coarse tiles of k_mr_interp (3-D tensor of float, box {2 CY, CX, 1}).
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int use_global, int c0, int c1, int c2, int bytes, float* out, int n) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        const CUtensorMap* m = use_global ? gmap : &tmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"((unsigned)__cvta_generic_to_shared(sm)), "l"(m), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled encode = (PFN_encodeTiled)fn;
    const int Md = 128, Nd = 128, NC = 7;
    float* d; cudaMalloc(&d, (size_t)NC * Nd * Md * 8);
    float* h = new float[(size_t)NC * Nd * Md * 2];
    for (size_t i = 0; i < (size_t)NC * Nd * Md * 2; ++i) h[i] = (float)i;
    cudaMemcpy(d, h, (size_t)NC * Nd * Md * 8, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 1 << 20);
    CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap));
    struct Cfg { int bx, by, c0, c1; } cfgs[] = {{64, 16, 0, 0}, {84, 26, 0, 0}, {84, 26, 118, 11}, {84, 26, 120, 11}, {88, 26, 116, 11}, {96, 26, 112, 11}};
    for (auto& c : cfgs)
        for (int use_global = 0; use_global < 2; ++use_global) {
            CUtensorMap tm;
            const cuuint64_t dims[3] = {2 * Md, Nd, NC};
            const cuuint64_t strides[2] = {Md * 8, (cuuint64_t)Nd * Md * 8};
            const cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, 1};
            const cuuint32_t es[3] = {1, 1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cudaMemcpy(gmap, &tm, sizeof(tm), cudaMemcpyHostToDevice);
            const int n = c.bx * c.by;
            k<<<1, 128, n * 4 + 256>>>(tm, gmap, use_global, c.c0, c.c1, 3, n * 4, out, n);
            cudaError_t e = cudaDeviceSynchronize();
            float v[2] = {-1, -1};
            if (e == cudaSuccess) { cudaMemcpy(v, out, 4, cudaMemcpyDeviceToHost); cudaMemcpy(v + 1, out + n - 1, 4, cudaMemcpyDeviceToHost); }
            const double want0 = ((double)3 * Nd + c.c1) * Md * 2 + c.c0, want1 = ((double)3 * Nd + c.c1 + c.by - 1) * Md * 2 + c.c0 + c.bx - 1;
            printf("box %dx%d at (%d,%d) desc=%s encode=%d run=%s first %.0f (want %.0f) last %.0f (want %.0f)\n", c.bx, c.by, c.c0, c.c1,
                   use_global ? "global" : "param", (int)r, cudaGetErrorString(e), v[0], want0, v[1], want1);
            if (e != cudaSuccess) { printf("context lost, stopping\n"); return 0; }
        }
    return 0;
}
