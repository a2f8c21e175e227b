// K2 row transforms, pipelined form (round 2): persistent CTAs, the next pair of rows arrives by ONE bulk copy
// (cp.async.bulk global -> shared, mbarrier completion; SASS UBLKCP) while the current pair is transformed.
//
// Same algorithm as k_dct2_rows_pow2 / k_idct2_rows_pow2 (unwrap.cu): Makhoul permutation, two rows per complex FFT,
// radix-2 / radix-4 head stage + radix-8 Stockham stages.  The twiddles w2, w4 come from squaring w1 and the Makhoul factors
// from one value per thread times constant rotations, so the results agree with those kernels to rounding (1e-16 relative on
// the transforms, 1e-13 on PCG iterates; tests/test_unwrap_gpu.py compares the two), not bit for bit.
// What changed is the data movement:
//   * the raw rows land in a staging buffer; the permutation (forward) / the DCT-III pre-twiddle (inverse) is folded into
//     the loads of the first FFT stage, so there is no separate "load + permute" pass and no thread ever waits on DRAM;
//   * as soon as every thread has read the staging buffer (first barrier of the FFT) one thread issues the bulk copy of the
//     CTA's NEXT pair of rows: DRAM latency and transfer overlap the whole transform (ncu of the old kernels: 43 % of the
//     stall samples were long_scoreboard on the first shared-memory store);
//   * the FFT buffer is padded by one complex every 8 (index p -> p + p/8): the head stage's stride-R stores were 4- / 8-way
//     bank conflicts (27 % of the shared-memory wavefronts);
//   * 34 n bytes of shared memory per CTA: three CTAs per SM at n = 2048 run in different phases of the transform.
// Reference semantics: scipy.fft.dctn / idctn as used by solvePoisson_precomped (pyGPA/phase_unwrap.py:95-103).
#pragma once
#include "fft_device.cuh"

namespace gpa {

__device__ __forceinline__ void dp_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dp_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "DPWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DPDONE_%=;\n\t"
        "bra DPWAIT_%=;\n\t"
        "DPDONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// one thread: announce `bytes` on the barrier and start the bulk copy global -> shared (both 16-byte aligned, bytes % 16 == 0)
__device__ __forceinline__ void dp_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the buffer's last generic-proxy accesses are ordered first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}

// 2-D tensor-map boxes (cp.async.bulk.tensor; SASS UTMALDG / UTMASTG): coordinates (c0 = innermost)
__device__ __forceinline__ void dp_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, unsigned long long* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dp_tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tmap), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dp_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ int dp_pad(int p) { return p + (p >> 3); }

// Twiddles: one compact shared-memory table per radix-8 stage, tws[off(NS) + k] = e^{-2 pi i k/(8 NS)}, k < NS — a thread of
// stage NS reads entry k = tid & (NS - 1), so a warp's reads are contiguous (conflict-free) — and w2 = w1^2, w4 = w2^2 by
// squaring.  (History: the global twiddle table missed the small L1 next to 3 x 70 KB of shared memory, 40 % of the stall
// samples on long_scoreboard; one table of the first octant indexed by k, 2k, 4k times the stage's step made 2- to 8-way
// bank conflicts and lifted the shared-memory pipe from 40 % to 56 % busy.)  sum of NS over the stages < n / 7 entries.
template <int LOGN>
__host__ __device__ constexpr int dp_tw_first() { return (LOGN % 3) ? 1 << (LOGN % 3) : 8; }
template <int LOGN>
__host__ __device__ constexpr int dp_tw_off(int NS) {
    int off = 0;
    for (int ns = dp_tw_first<LOGN>(); ns < NS; ns <<= 3) off += ns;
    return off;
}
template <int LOGN>
__host__ __device__ constexpr int dp_tw_entries() { return dp_tw_off<LOGN>(1 << LOGN); }

// all threads of the CTA: fill the stage tables from the global table tw[t] = e^{-2 pi i t/n}
template <int LOGN>
__device__ __forceinline__ void dp_tw_fill(double2* __restrict__ tws, const double2* __restrict__ tw, int tid, int nthr) {
    constexpr int n = 1 << LOGN;
    int off = 0;
#pragma unroll
    for (int ns = dp_tw_first<LOGN>(); ns < n; ns <<= 3) {
        const int step = n / (8 * ns);
        for (int k = tid; k < ns; k += nthr) tws[off + k] = tw[k * step];
        off += ns;
    }
}

// e^{-i pi m/16}, m = 0..7: mk[tid + m n/8] = mk[tid] e^{-i pi m/16} (mk[k] = e^{-i pi k/2n}), so a thread keeps ONE Makhoul
// factor in registers for all its eight positions and all the rows it transforms.
__device__ __forceinline__ double2 dp_rot16(const int m) {
    constexpr double c1 = 0.98078528040323044912618223613424, s1 = 0.19509032201612826784828486847702;
    constexpr double c2 = 0.92387953251128675612818318939679, s2 = 0.38268343236508977172845998403040;
    constexpr double c3 = 0.83146961230254523707878837761791, s3 = 0.55557023301960222474283081394853;
    constexpr double c4 = 0.70710678118654752440084436210485;
    switch (m) {
        case 0: return make_double2(1.0, 0.0);
        case 1: return make_double2(c1, -s1);
        case 2: return make_double2(c2, -s2);
        case 3: return make_double2(c3, -s3);
        case 4: return make_double2(c4, -c4);
        case 5: return make_double2(s3, -c3);
        case 6: return make_double2(s2, -c2);
        default: return make_double2(s1, -c1);
    }
}

// radix-8 Stockham stages NS, 8 NS, ... of an n = 2^LOGN point FFT held in the padded buffer; n / 8 threads (tid = index in
// the group), one butterfly per thread.  FIRST: the inputs come from get(p) instead of the buffer (no twiddles at NS = 1).
// idx(p) = position of element p in the buffer (dp_pad(p) for one FFT per CTA, 2 dp_pad(p) + g for two interleaved FFTs).
template <int LOGN, int NS, bool FIRST, typename Get, typename Sync, typename Hook, typename Idx>
__device__ __forceinline__ void dp_stage8(double2* __restrict__ buf, const double2* __restrict__ t8, const int tid, Get get,
                                          Sync sync, Hook after_first_read, Idx idx) {
    constexpr int n = 1 << LOGN, e = n >> 3;
    if constexpr (NS < n) {
        double2 u[8];
        const int j = tid, k = j & (NS - 1);
        if constexpr (FIRST) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = get(i);        // position tid + i n/8
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = buf[idx(j + i * e)];
        }
        if constexpr (NS > 1) {
            const double2 w1 = t8[dp_tw_off<LOGN>(NS) + k];          // e^{-2 pi i k/(8 NS)}; w_q = w1^q
            const double2 w2 = zmul(w1, w1), w4 = zmul(w2, w2);
            const double2 w3 = zmul(w1, w2), w5 = zmul(w1, w4), w6 = zmul(w2, w4);
            const double2 w7 = zmul(w3, w4);
            u[1] = zmul(u[1], w1);
            u[2] = zmul(u[2], w2);
            u[3] = zmul(u[3], w3);
            u[4] = zmul(u[4], w4);
            u[5] = zmul(u[5], w5);
            u[6] = zmul(u[6], w6);
            u[7] = zmul(u[7], w7);
        }
        sync();
        if constexpr (FIRST) after_first_read();
        const int j0 = ((j - k) << 3) + k;
        dft8(u);
#pragma unroll
        for (int i = 0; i < 8; ++i) buf[idx(j0 + i * NS)] = u[i];
        sync();
        dp_stage8<LOGN, NS * 8, false>(buf, t8, tid, get, sync, after_first_read, idx);
    }
}

// buf (padded) <- FFT_n of the input sequence; n / 8 threads; get(m) returns the thread's input at position tid + m n/8,
// m = 0..7 (every head stage reads exactly these eight).  after_first_read() runs once every thread has consumed
// its inputs (the staging buffer may be refilled from then on).  On return buf is visible to all threads of the group.
template <int LOGN, typename Get, typename Sync, typename Hook, typename Idx>
__device__ __forceinline__ void dp_fft(double2* __restrict__ buf, const double2* __restrict__ t8, const int tid, Get get,
                                       Sync sync, Hook after_first_read, Idx idx) {
    constexpr int n = 1 << LOGN, T = n >> 3;
    constexpr int R0 = 1 << (LOGN % 3);
    if constexpr (R0 == 2) {                             // one radix-2 stage (ns = 1: twiddles are 1), 4 butterflies per thread
        constexpr int half = n >> 1;
        double2 a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] = get(q);               // j
            b[q] = get(q + 4);           // j + n/2
        }
        sync();
        after_first_read();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = tid + q * T;
            buf[idx(2 * j)] = zadd(a[q], b[q]);
            buf[idx(2 * j + 1)] = zsub(a[q], b[q]);
        }
        sync();
        dp_stage8<LOGN, 2, false>(buf, t8, tid, get, sync, after_first_read, idx);
    } else if constexpr (R0 == 4) {                      // one radix-4 stage, 2 butterflies per thread
        constexpr int quarter = n >> 2;
        double2 v[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = tid + q * T;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[q][i] = get(q + 2 * i);      // j + i n/4
        }
        sync();
        after_first_read();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = tid + q * T;
            dft4(v[q][0], v[q][1], v[q][2], v[q][3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) buf[idx(4 * j + i)] = v[q][i];
        }
        sync();
        dp_stage8<LOGN, 4, false>(buf, t8, tid, get, sync, after_first_read, idx);
    } else {
        dp_stage8<LOGN, 1, true>(buf, t8, tid, get, sync, after_first_read, idx);
    }
}

// staging (two raw rows) | padded FFT buffer | first-octant twiddles
// column strips: two interleaved padded FFT buffers (the raw strip of 4 columns lands in the same memory) | twiddles
constexpr size_t dp_cols_smem_bytes(int n) {
    return (size_t)2 * (n + n / 8) * sizeof(double2) + (size_t)(n / 7 + 1) * sizeof(double2);
}
constexpr size_t dp_rows_smem_bytes(int n) {
    return (size_t)2 * n * sizeof(double) + (size_t)(n + n / 8) * sizeof(double2) + (size_t)(n / 7 + 1) * sizeof(double2);
}

}  // namespace gpa
