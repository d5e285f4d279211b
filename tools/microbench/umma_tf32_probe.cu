// Development probe (not part of the library, not run by the tests): one CTA
// computes C[128 x N] = A[128 x K] * B[N x K]^T with tcgen05.mma kind::tf32,
// accumulators in TMEM, operands K-major in shared memory under the 128-byte
// swizzle.  It exists to pin, on the device, the three encodings the float
// contraction kernels need before they move from mma.sync to tcgen05:
//   (1) the shared-memory matrix descriptor of a K-major SWIZZLE_128B tile,
//   (2) the instruction descriptor of kind::tf32 with FP32 accumulation,
//   (3) the error-compensated 3xTF32 product (a = a_hi + a_lo split in shared
//       memory, D += a_lo b_hi + a_hi b_lo + a_hi b_hi) at tensor rate.
// Field positions follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor,
// UMMA::InstrDescriptor) of the CUTLASS tree vendored in this image; the PTX
// follows /opt/skills/guides/blackwell_cuda_programming.md section 1-2.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -shared -Xcompiler -fPIC -o umma_tf32_probe.so umma_tf32_probe.cu
//   python tools/microbench/umma_probe.py        (needs a B200)
#include <cstdint>
#include <cuda_runtime.h>

namespace
{
constexpr int M  = 128; // UMMA M (one CTA, cta_group::1)
constexpr int KB = 32;  // floats per 128-byte swizzle row = one k-block
constexpr int UK = 8;   // K of one kind::tf32 instruction (32 bytes)

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// byte offset of element (row, k) in a K-major tile of 32 floats per row laid
// out as 8-row x 128-byte atoms (1024 B each, atoms consecutive along rows) with
// Swizzle<3,4,3>: the 16-byte chunk index is XORed with the row within the atom
__device__ __forceinline__ uint32_t swz128(int row, int k)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2)));
}

// UMMA::SmemDescriptor: start address [0,14) (>>4), leading byte offset [16,30)
// (>>4; 1 for swizzled K-major), stride byte offset [32,46) (>>4; 1024 B between
// 8-row atoms), version [46,48) = 1 on sm_100, base offset [49,52) = 0 (tile
// 1024-byte aligned), layout type [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// UMMA::InstrDescriptor: c_format [4,6) = 1 (F32), a_format [7,10) = b_format
// [10,13) = 2 (TF32), a_major [15] = b_major [16] = 0 (K), n_dim [17,23) = N>>3,
// m_dim [24,29) = M>>4
__device__ __forceinline__ uint32_t make_instr_desc(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc,
    uint32_t accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_c),
                 "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}

__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "W:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra D;\n\t"
                 "bra W;\n\t"
                 "D:\n\t"
                 "}\n" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
}

// A: [128][K] floats, B: [n][K] floats (both K contiguous), C: [128][n] floats.
// K a multiple of 32; n a multiple of 32, 32 <= n <= 256.  split = 0: operands are
// used as they are (pass TF32-representable values for an exact check); split =
// 1: 3xTF32.  flush > 0: every `flush` k-blocks the TMEM accumulator is drained
// into a double accumulator in registers (the long-K scheme of the contractions).
extern "C" __global__ void __launch_bounds__(128, 1)
    k_umma_tf32_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
        int n, int K, int split, int ncols)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t* a_hi = smem;                   // 128 rows x 128 B
    uint8_t* a_lo = a_hi + M * 128;
    uint8_t* b_hi = a_lo + M * 128;         // n rows x 128 B
    uint8_t* b_lo = b_hi + 256 * 128;

    if (t == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0)
    {
        // ncols: power of two >= 32 covering the n accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)),
                     "r"((uint32_t)ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = tmem_base_s;
    const uint32_t idesc  = make_instr_desc(n);
    const uint32_t bar    = smem_u32(&mbar);

    uint32_t phase = 0;
    const int nkb  = K / KB;
    for (int kb = 0; kb < nkb; kb++)
    {
        // stage one k-block: rows of 32 floats, swizzled; with the split, the TF32
        // head goes to *_hi and the TF32-rounded remainder to *_lo
        for (int i = t; i < M * KB; i += 128)
        {
            const int row = i / KB, k = i % KB;
            const float a = A[(size_t)row * K + kb * KB + k];
            const uint32_t o = swz128(row, k);
            if (split)
            {
                const float h                    = to_tf32(a);
                *(float*)(a_hi + o)              = h;
                *(float*)(a_lo + o)              = to_tf32(a - h);
            }
            else
                *(float*)(a_hi + o) = a;
        }
        for (int i = t; i < n * KB; i += 128)
        {
            const int row = i / KB, k = i % KB;
            const float b = B[(size_t)row * K + kb * KB + k];
            const uint32_t o = swz128(row, k);
            if (split)
            {
                const float h       = to_tf32(b);
                *(float*)(b_hi + o) = h;
                *(float*)(b_lo + o) = to_tf32(b - h);
            }
            else
                *(float*)(b_hi + o) = b;
        }
        // generic-proxy writes -> visible to the async proxy (the tensor core)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (t == 0)
        {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dah = make_smem_desc(smem_u32(a_hi)), dal = make_smem_desc(smem_u32(a_lo));
            const uint64_t dbh = make_smem_desc(smem_u32(b_hi)), dbl = make_smem_desc(smem_u32(b_lo));
#pragma unroll
            for (int ks = 0; ks < KB / UK; ks++)
            {
                // 32 bytes further along K inside the swizzle row: +2 in the
                // 16-byte units of the start-address field
                const uint64_t adv = (uint64_t)((ks * UK * 4) >> 4);
                const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                if (split)
                {
                    umma_tf32(tmem_c, dal + adv, dbh + adv, idesc, acc);
                    umma_tf32(tmem_c, dah + adv, dbl + adv, idesc, 1u);
                    umma_tf32(tmem_c, dah + adv, dbh + adv, idesc, 1u);
                }
                else
                    umma_tf32(tmem_c, dah + adv, dbh + adv, idesc, acc);
            }
            // arrives on the mbarrier when every MMA issued so far has completed
            // (and has therefore finished reading shared memory)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                         : "memory");
        }
        mbar_wait(bar, phase);
        phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows of C; 32 columns at a time
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < n; c0 += 32)
    {
        uint32_t v[32];
        const uint32_t taddr = tmem_c + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++)
            C[(size_t)row * n + c0 + j] = __uint_as_float(v[j]);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "r"((uint32_t)ncols)
                     : "memory");
}

// Host entry for ctypes: device pointers in, 0 on success or the CUDA error code.
extern "C" int umma_tf32_probe(const float* A, const float* B, float* C, int n, int K, int split)
{
    if (n < 32 || n > 256 || n % 32 || K < 32 || K % 32) return -1;
    int ncols = 32;
    while (ncols < n) ncols *= 2;
    const size_t smem = (size_t)(2 * M * 128 + 2 * 256 * 128) + 1024;
    cudaError_t e     = cudaFuncSetAttribute(
        k_umma_tf32_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_umma_tf32_probe<<<1, 128, smem>>>(A, B, C, n, K, split, ncols);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}
