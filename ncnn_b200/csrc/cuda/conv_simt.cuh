// conv_simt.cuh -- CUDA-core (FFMA, fp32 accumulate) implicit-GEMM convolution.
// This is the strict-fp32 path (parity target <= 1e-5 normalised vs the reference's fp32 CPU path,
// src/layer/convolution.cpp:113-184) and the fallback for shapes the tensor-core path does not take.
//   M = n * outh * outw (output pixels), N = outch, K = kh * kw * inch  (k = (ky*kw + kx)*inch + ci)
#pragma once
#include "common.cuh"

namespace ncnn_cuda {

struct ConvGeom
{
    int inch, outch;
    int kw, kh, dw, dh, sw, sh;
    int pad_left, pad_top;
    float pad_value;
    int inw, inh, outw, outh, n;
    int in_cpitch, out_cpitch, res_cpitch;
    long long in_nstep, out_nstep, res_nstep;
    int K;     // real reduction length
    int wp_ld; // leading dimension (padded outch) of the packed weights [Kpad][wp_ld]
    int act_type;
    float act_p0, act_p1;
};

// BM x BN output tile per CTA, BK = 16, 256 threads, each thread (BM/16) x (BN/16) outputs
template<typename T, int BM, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const T* __restrict__ in, const float* __restrict__ wp, const float* __restrict__ bias,
                                                        const T* __restrict__ residual, T* __restrict__ out, ConvGeom g)
{
    constexpr int BK = 16;
    constexpr int TM = BM / 16;
    constexpr int TN = BN / 16;
    constexpr int AROWS = BM / 16; // rows of A each thread stages per k-tile

    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15;  // N direction
    const int ty = tid >> 4;  // M direction
    const long long M = (long long)g.n * g.outh * g.outw;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A staging: this thread always fetches k-lane `ak` of rows ar + 16*i
    const int ak = tid & 15;
    const int ar = tid >> 4;
    long long a_base[AROWS]; // offset of (n, 0, 0, 0) for the row, or -1 if the row is past M
    int a_iy0[AROWS], a_ix0[AROWS];
    const int opix = g.outh * g.outw;
#pragma unroll
    for (int i = 0; i < AROWS; i++)
    {
        long long m = m0 + ar + 16 * i;
        if (m < M)
        {
            int b = (int)(m / opix);
            int p = (int)(m - (long long)b * opix);
            int oy = p / g.outw;
            int ox = p - oy * g.outw;
            a_base[i] = (long long)b * g.in_nstep;
            a_iy0[i] = oy * g.sh - g.pad_top;
            a_ix0[i] = ox * g.sw - g.pad_left;
        }
        else
        {
            a_base[i] = -1;
            a_iy0[i] = 0;
            a_ix0[i] = 0;
        }
    }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    const int ktiles = (g.K + BK - 1) / BK;
    for (int kt = 0; kt < ktiles; kt++)
    {
        // ---- stage A (gathered im2col column) and B (packed weights)
        {
            int k = kt * BK + ak;
            bool kvalid = k < g.K;
            int ci = 0, ky = 0, kx = 0;
            if (kvalid)
            {
                int tap = k / g.inch;
                ci = k - tap * g.inch;
                ky = tap / g.kw;
                kx = tap - ky * g.kw;
            }
#pragma unroll
            for (int i = 0; i < AROWS; i++)
            {
                float v = 0.f;
                if (kvalid && a_base[i] >= 0)
                {
                    int iy = a_iy0[i] + ky * g.dh;
                    int ix = a_ix0[i] + kx * g.dw;
                    if (iy >= 0 && iy < g.inh && ix >= 0 && ix < g.inw)
                        v = to_f32(in[a_base[i] + ((long long)iy * g.inw + ix) * g.in_cpitch + ci]);
                    else
                        v = g.pad_value;
                }
                As[ak][ar + 16 * i] = v;
            }
            // B: BK x BN floats, rows contiguous in wp
            for (int e = tid; e < BK * BN / 4; e += 256)
            {
                int r = e / (BN / 4);
                int c4 = e - r * (BN / 4);
                const float4 w4 = *reinterpret_cast<const float4*>(wp + (long long)(kt * BK + r) * g.wp_ld + n0 + c4 * 4);
                *reinterpret_cast<float4*>(&Bs[r][c4 * 4]) = w4;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++)
        {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, residual, activation
#pragma unroll
    for (int i = 0; i < TM; i++)
    {
        long long m = m0 + ty * TM + i;
        if (m >= M) continue;
        int b = (int)(m / opix);
        long long p = m - (long long)b * opix;
        T* orow = out + (long long)b * g.out_nstep + p * g.out_cpitch;
        const T* rrow = residual ? residual + (long long)b * g.res_nstep + p * g.res_cpitch : 0;
#pragma unroll
        for (int j = 0; j < TN; j++)
        {
            int oc = n0 + tx * TN + j;
            if (oc >= g.outch) continue;
            float v = acc[i][j];
            if (bias) v += bias[oc];
            if (rrow) v += to_f32(rrow[oc]);
            v = apply_activation(v, g.act_type, g.act_p0, g.act_p1);
            orow[oc] = from_f32<T>(v);
        }
    }
}

template<typename T>
static int launch_conv_simt(const T* in, const float* wp, const float* bias, const T* residual, T* out, const ConvGeom& g, cudaStream_t stream)
{
    long long M = (long long)g.n * g.outh * g.outw;
    if (M == 0) return 0;
    if (g.outch > 64)
    {
        dim3 grid(ceil_div(M, 128), ceil_div(g.outch, 128));
        conv_simt_kernel<T, 128, 128><<<grid, 256, 0, stream>>>(in, wp, bias, residual, out, g);
    }
    else if (M >= 4096)
    {
        dim3 grid(ceil_div(M, 128), ceil_div(g.outch, 64));
        conv_simt_kernel<T, 128, 64><<<grid, 256, 0, stream>>>(in, wp, bias, residual, out, g);
    }
    else
    {
        dim3 grid(ceil_div(M, 64), ceil_div(g.outch, 64));
        conv_simt_kernel<T, 64, 64><<<grid, 256, 0, stream>>>(in, wp, bias, residual, out, g);
    }
    NC_LAUNCH_CHECK();
    return 0;
}

} // namespace ncnn_cuda
