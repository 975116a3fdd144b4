// fp32 implicit-GEMM kernel (CUDA-core FFMA path).  See gemm.cuh for the operand model.
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, register-prefetched double buffering.
#include "gemm.cuh"

#define BM 128
#define BN 64
#define BK 16

template <int NEL, int ROWS, bool CK>
struct TileLoader {
    // CK  : k = t % 16, row_i = t / 16 + 16 * i            (consecutive threads walk along k)
    // !CK : row = t % ROWS, k_i = t / ROWS + (256/ROWS) * i (consecutive threads walk along rows)
    __device__ __forceinline__ static int row(int t, int i) { return CK ? (t >> 4) + 16 * i : t % ROWS; }
    __device__ __forceinline__ static int kk(int t, int i) { return CK ? (t & 15) : t / ROWS + (256 / ROWS) * i; }
};

template <int NEL, int ROWS, bool CK>
__device__ __forceinline__ void fetch_tile(const GOperand& o, int r0, int R, int k0, int kend, int t, const SpIdx* pre,
                                           float* regs) {
    using L = TileLoader<NEL, ROWS, CK>;
    if (o.mode == OPM_STRIDED) {
#pragma unroll
        for (int i = 0; i < NEL; i++) {
            int r = r0 + L::row(t, i), k = k0 + L::kk(t, i);
            regs[i] = (r < R && k < kend) ? __ldg(o.p + (long long)r * o.s_a + (long long)k * o.s_b) : 0.f;
        }
    } else if (!o.swap) {
#pragma unroll
        for (int i = 0; i < NEL; i++) {
            int r = r0 + L::row(t, i), k = k0 + L::kk(t, i);
            regs[i] = (r < R && k < kend) ? gather_elem(o, pre[CK ? i : 0], k) : 0.f;
        }
    } else {
        SpIdx s;
        if (CK) s = decode_sp(o.X, o.Y, o.Z, min(k0 + L::kk(t, 0), kend - 1));
#pragma unroll
        for (int i = 0; i < NEL; i++) {
            int r = r0 + L::row(t, i), k = k0 + L::kk(t, i);
            if (r < R && k < kend) {
                if (!CK) s = decode_sp(o.X, o.Y, o.Z, k);
                regs[i] = gather_elem(o, s, r);
            } else {
                regs[i] = 0.f;
            }
        }
    }
}

template <bool A_CK, bool B_CK>
__global__ void __launch_bounds__(256) gemm_kernel(const __grid_constant__ GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * p.ksplit, kend = min(p.K, kbeg + p.ksplit);
    using LA = TileLoader<8, BM, A_CK>;
    using LB = TileLoader<4, BN, B_CK>;

    SpIdx preA[A_CK ? 8 : 1], preB[B_CK ? 4 : 1];
    if (p.A.mode != OPM_STRIDED && !p.A.swap) {
#pragma unroll
        for (int i = 0; i < (A_CK ? 8 : 1); i++) preA[i] = decode_sp(p.A.X, p.A.Y, p.A.Z, min(m0 + LA::row(t, i), p.M - 1));
    }
    if (p.B.mode != OPM_STRIDED && !p.B.swap) {
#pragma unroll
        for (int i = 0; i < (B_CK ? 4 : 1); i++) preB[i] = decode_sp(p.B.X, p.B.Y, p.B.Z, min(n0 + LB::row(t, i), p.N - 1));
    }

    float ra[8], rb[4];
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    const int ty = t >> 4, tx = t & 15;
    int buf = 0;
    if (kbeg < kend) {
        fetch_tile<8, BM, A_CK>(p.A, m0, p.M, kbeg, kend, t, preA, ra);
        fetch_tile<4, BN, B_CK>(p.B, n0, p.N, kbeg, kend, t, preB, rb);
#pragma unroll
        for (int i = 0; i < 8; i++) As[0][LA::kk(t, i)][LA::row(t, i)] = ra[i];
#pragma unroll
        for (int i = 0; i < 4; i++) Bs[0][LB::kk(t, i)][LB::row(t, i)] = rb[i];
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) {
            fetch_tile<8, BM, A_CK>(p.A, m0, p.M, k0 + BK, kend, t, preA, ra);
            fetch_tile<4, BN, B_CK>(p.B, n0, p.N, k0 + BK, kend, t, preB, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
#pragma unroll
            for (int i = 0; i < 8; i++) As[buf ^ 1][LA::kk(t, i)][LA::row(t, i)] = ra[i];
#pragma unroll
            for (int i = 0; i < 4; i++) Bs[buf ^ 1][LB::kk(t, i)][LB::row(t, i)] = rb[i];
        }
        __syncthreads();
        buf ^= 1;
    }

    // ------------------------------------------------------------------ epilogue
    const GEpilogue& e = p.E;
    const int k3 = e.ks * e.ks * e.ks;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
        SpIdx s;
        if (e.flags & EPI_D2S) s = decode_sp(e.X, e.Y, e.Z, m);
        float rs = 1.f;
        if ((e.flags & EPI_RESID) && e.row_scale) rs = e.row_scale[m / e.rows_per_scale];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            long long idx = (e.flags & EPI_D2S) ? d2s_addr(e.X, e.Y, e.Z, e.ld, e.ks, s, n) : (long long)m * e.ldc + n;
            if (e.flags & EPI_BIAS) v += e.bias[(e.flags & EPI_D2S) ? n / k3 : n];
            if (e.flags & EPI_GELU) {
                e.aux[idx] = v;
                v = gelu_erf(v);
            }
            if (e.flags & EPI_GELU_GRAD) v *= gelu_erf_grad(e.aux[idx]);
            if (e.flags & EPI_RESID) v = e.resid[idx] + rs * v;
            if (e.flags & EPI_ATOMIC)
                atomicAdd(e.out + idx, v);
            else if (e.flags & EPI_ACCUM)
                e.out[idx] += v;
            else
                e.out[idx] = v;
        }
    }
}

static bool contig_k(const GOperand& o) {
    if (o.mode == OPM_STRIDED) return o.s_b == 1 || o.s_a != 1;
    return !o.swap;
}

int nmae_gemm_launch(const GemmParams& p, cudaStream_t stream) {
    NMAE_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
    int splits = cdiv(p.K, p.ksplit);
    NMAE_CHECK_ARG(splits == 1 || (p.E.flags & EPI_ATOMIC), "gemm: split-K requires EPI_ATOMIC");
    dim3 grid(cdiv(p.M, BM), cdiv(p.N, BN), splits);
    NMAE_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm: grid too large");
    bool a = contig_k(p.A), b = contig_k(p.B);
    if (a && b)
        gemm_kernel<true, true><<<grid, 256, 0, stream>>>(p);
    else if (a && !b)
        gemm_kernel<true, false><<<grid, 256, 0, stream>>>(p);
    else if (!a && b)
        gemm_kernel<false, true><<<grid, 256, 0, stream>>>(p);
    else
        gemm_kernel<false, false><<<grid, 256, 0, stream>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
