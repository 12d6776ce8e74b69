// pe_sell.cu -- sliced-ELL (SELL-32) storage and the thread-per-row streaming kernels built
// on it: SpMV / residual / prolongation (K1/K2/K7) and one colour of multicolour
// (l1-)Gauss-Seidel (K4).
//
// Layout: rows are grouped in slices of 32 (one warp); slice s stores w_s = max row length
// of its rows, column-major: entry q of lane l sits at (soff[s] + q) * 32 + l.  A warp-wide
// load of J is one 128-byte line and of A two lines (always fully used), and because
// consecutive rows of a (locally) structured operator couple to consecutive columns, the
// x-gather of one q touches ~8-10 sectors instead of the ~25 of the CSR layout.  Short rows
// are padded with (col = last valid column of the row, val = 0).
//
// The SELL copy is what the kernels stream; the CSR block stays the interchange format of
// the C ABI (upload/download/SpGEMM).  Algorithmic bytes are counted on the CSR figures
// (12 B per true non-zero), see DESIGN.md.
#include "pe_core.cuh"
#include "pe_stream.cuh"
#include <cub/cub.cuh>

// ---------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------
__device__ __forceinline__ int sell_src_row(const int *rowmap, int k, int nrows_src)
{
    if (rowmap) return rowmap[k];
    return k < nrows_src ? k : -1;
}

// one warp per slice: width = max over its rows of (diag + offd) row length
__global__ void k_sell_widths(int nslices, int nrows_src, const int *__restrict__ rowmap,
                              const int *__restrict__ dI, const int *__restrict__ oI, int *w)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = sell_src_row(rowmap, s * 32 + lane, nrows_src);
    int len = 0;
    if (r >= 0) { len = dI[r + 1] - dI[r]; if (oI) len += oI[r + 1] - oI[r]; }
    len = __reduce_max_sync(0xffffffffu, len);
    if (lane == 0) { w[s + 1] = len; if (s == 0) w[0] = 0; }
}

__global__ void k_sell_fill(int nslices, int nrows_src, const int *__restrict__ rowmap,
                            const int *__restrict__ colpos, int ext_base,
                            const int *__restrict__ dI, const int *__restrict__ dJ, const double *__restrict__ dA,
                            const int *__restrict__ oI, const int *__restrict__ oJ, const double *__restrict__ oA,
                            const int *__restrict__ soff, int *J, double *A)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = sell_src_row(rowmap, s * 32 + lane, nrows_src);
    const int o0 = soff[s], w = soff[s + 1] - o0;
    int64_t p = (int64_t)o0 * 32 + lane;
    int q = 0, last = 0;
    if (r >= 0)
    {
        for (int t = dI[r]; t < dI[r + 1]; ++t, ++q, p += 32)
        {
            last = colpos ? colpos[dJ[t]] : dJ[t];
            J[p] = last; A[p] = dA[t];
        }
        if (oI) for (int t = oI[r]; t < oI[r + 1]; ++t, ++q, p += 32)
        {
            last = ext_base + oJ[t];
            J[p] = last; A[p] = oA[t];
        }
    }
    for (; q < w; ++q, p += 32) { J[p] = last; A[p] = 0.0; }
}

void pe_sell_free(DevSELL &m)
{
    if (m.soff) cudaFree(m.soff);
    if (m.J) cudaFree(m.J);
    if (m.A) cudaFree(m.A);
    m = DevSELL();
}

// rowmap_d: nslices*32 source rows (-1 = padding row) or null for the identity over diag.nrows;
// colpos_d: new index of every diag column or null; ghost column j becomes ext_base + j.
int pe_sell_build(pe_ctx *ctx, const DevCSR &diag, const DevCSR *offd, const int32_t *rowmap_d, int32_t nslices,
                  const int32_t *colpos_d, int32_t ext_base, DevSELL &out)
{
    cudaStream_t st = ctx->stream;
    if (!rowmap_d) nslices = (diag.nrows + 31) / 32;
    out = DevSELL();
    out.nslices = nslices;
    out.nrows = diag.nrows;
    PE_CUDA(cudaMalloc(&out.soff, sizeof(int32_t) * (size_t)(nslices + 1)));
    if (nslices == 0) { PE_CUDA(cudaMemsetAsync(out.soff, 0, sizeof(int32_t), st)); return 0; }
    const int *oI = (offd && offd->nnz > 0) ? offd->I : nullptr;
    const int grid = pe_grid_for((int64_t)nslices * 32, 256);
    k_sell_widths<<<grid, 256, 0, st>>>(nslices, diag.nrows, rowmap_d, diag.I, oI, out.soff);
    PE_LAUNCHED(ctx);
    void *tmp = nullptr; size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, out.soff, out.soff, nslices + 1, st);
    PE_CUDA(cudaMalloc(&tmp, tb));
    PE_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, out.soff, out.soff, nslices + 1, st));
    ctx->launches++;
    int32_t total = 0;
    PE_CUDA(cudaMemcpyAsync(&total, out.soff + nslices, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PE_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    out.nstored = (int64_t)total * 32;
    PE_CUDA(cudaMalloc(&out.J, sizeof(int32_t) * (size_t)(out.nstored > 0 ? out.nstored : 1)));
    PE_CUDA(cudaMalloc(&out.A, sizeof(double) * (size_t)(out.nstored > 0 ? out.nstored : 1)));
    if (out.nstored > 0)
    {
        k_sell_fill<<<grid, 256, 0, st>>>(nslices, diag.nrows, rowmap_d, colpos_d, ext_base, diag.I, diag.J, diag.A,
                                          oI, oI ? offd->J : nullptr, oI ? offd->A : nullptr, out.soff, out.J, out.A);
        PE_LAUNCHED(ctx);
    }
    return 0;
}

// Lazily attach a SELL copy to a CSR block used by SpMV.  Rejected (state -1) when padding
// would add more than 25 % to the stream.
int pe_sell_for_spmv(pe_ctx *ctx, DevCSR &m)
{
    if (m.sell_state != 0) return 0;
    m.sell_state = -1;
    if (m.nnz == 0 || m.nrows == 0) return 0;
    DevSELL *s = new DevSELL();
    int rc = pe_sell_build(ctx, m, nullptr, nullptr, 0, nullptr, 0, *s);
    if (rc) { delete s; return rc; }
    if ((double)s->nstored > 1.25 * (double)m.nnz) { pe_sell_free(*s); delete s; return 0; }
    m.sell = s;
    m.sell_state = 1;
    return 0;
}

// ---------------------------------------------------------------------------
// SpMV: yout = alpha * A x + beta * yin   (thread per row, warp per slice)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sell_spmv(int nslices, int nrows, const int *__restrict__ soff, const int *__restrict__ J,
            const double *__restrict__ A, const double *__restrict__ x, double alpha, double beta,
            const double *yin, double *yout)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int o0 = soff[s], w = soff[s + 1] - o0;
    const int *j = J + (int64_t)o0 * 32 + lane;
    const double *a = A + (int64_t)o0 * 32 + lane;
    const uint64_t pol = l2_evict_first_policy();
    double acc = 0.0;
    int q = 0;
    for (; q + 4 <= w; q += 4)
    {
        const int c0 = ld_stream_s32(j + (q + 0) * 32, pol), c1 = ld_stream_s32(j + (q + 1) * 32, pol);
        const int c2 = ld_stream_s32(j + (q + 2) * 32, pol), c3 = ld_stream_s32(j + (q + 3) * 32, pol);
        const double a0 = ld_stream_f64(a + (q + 0) * 32, pol), a1 = ld_stream_f64(a + (q + 1) * 32, pol);
        const double a2 = ld_stream_f64(a + (q + 2) * 32, pol), a3 = ld_stream_f64(a + (q + 3) * 32, pol);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        acc += a0 * x0; acc += a1 * x1; acc += a2 * x2; acc += a3 * x3;
    }
    for (; q < w; ++q) acc += ld_stream_f64(a + q * 32, pol) * __ldg(x + ld_stream_s32(j + q * 32, pol));
    const int row = s * 32 + lane;
    if (row < nrows)
    {
        double v = alpha * acc;
        if (beta != 0.0) v += beta * yin[row];
        yout[row] = v;
    }
}

int pe_launch_sell_spmv(pe_ctx *ctx, const DevSELL &S, double alpha, const double *x, double beta,
                        const double *yin, double *yout)
{
    if (S.nslices == 0) return 0;
    if (ctx->rec)
    {
        PeOp o = pe_op(PE_OP_SELL_SPMV); o.i0 = S.nslices; o.i1 = S.nrows; o.a = alpha; o.b = beta;
        o.p[0] = S.soff; o.p[1] = S.J; o.p[2] = S.A; o.p[3] = x; o.p[4] = yin; o.p[5] = yout;
        pe_rec_push(ctx, o, -1.0);
        return 0;
    }
    k_sell_spmv<<<pe_grid_for((int64_t)S.nslices * 32, 256), 256, 0, ctx->stream>>>(S.nslices, S.nrows, S.soff, S.J, S.A,
                                                                                   x, alpha, beta, yin, yout);
    PE_LAUNCHED(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// One colour of (l1-)Gauss-Seidel on the colour-ordered SELL copy: slices [s0, s1), in-place
// update of u (colour-ordered numbering, padded rows have l1 = 0 and are skipped).  Columns
// >= ext_base address the halo.  Rows of one colour are mutually independent, so no thread
// reads a value another thread of this launch writes.
// ---------------------------------------------------------------------------
template <bool OFFD>
__global__ void __launch_bounds__(256)
k_sell_gs(int s0, int s1, const int *__restrict__ soff, const int *__restrict__ J, const double *__restrict__ A,
          int ext_base, const double *__restrict__ f, double *u, const double *__restrict__ uext,
          const double *__restrict__ l1)
{
    const int s = s0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= s1) return;
    const int o0 = soff[s], w = soff[s + 1] - o0;
    const int *j = J + (int64_t)o0 * 32 + lane;
    const double *a = A + (int64_t)o0 * 32 + lane;
    const int row = s * 32 + lane;
    const double d = l1[row], fr = f[row];
    const uint64_t pol = l2_evict_first_policy();
    double acc = 0.0;
    int q = 0;
#define PE_GATHER(c) ((OFFD && (c) >= ext_base) ? uext[(c) - ext_base] : u[(c)])
    // groups of 4 entries, software-pipelined: the (col,val) loads of group g+1 are in flight
    // while the gathers of group g are outstanding
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (w >= 4)
    {
        c0 = ld_stream_s32(j, pol); c1 = ld_stream_s32(j + 32, pol); c2 = ld_stream_s32(j + 64, pol); c3 = ld_stream_s32(j + 96, pol);
        a0 = ld_stream_f64(a, pol); a1 = ld_stream_f64(a + 32, pol); a2 = ld_stream_f64(a + 64, pol); a3 = ld_stream_f64(a + 96, pol);
    }
    for (; q + 4 <= w; q += 4)
    {
        const double u0 = PE_GATHER(c0), u1 = PE_GATHER(c1), u2 = PE_GATHER(c2), u3 = PE_GATHER(c3);
        const double b0 = a0, b1 = a1, b2 = a2, b3 = a3;
        if (q + 8 <= w)
        {
            const int *jn = j + (q + 4) * 32; const double *an = a + (q + 4) * 32;
            c0 = ld_stream_s32(jn, pol); c1 = ld_stream_s32(jn + 32, pol); c2 = ld_stream_s32(jn + 64, pol); c3 = ld_stream_s32(jn + 96, pol);
            a0 = ld_stream_f64(an, pol); a1 = ld_stream_f64(an + 32, pol); a2 = ld_stream_f64(an + 64, pol); a3 = ld_stream_f64(an + 96, pol);
        }
        acc += b0 * u0; acc += b1 * u1; acc += b2 * u2; acc += b3 * u3;
    }
    for (; q < w; ++q) { const int c = ld_stream_s32(j + q * 32, pol); acc += ld_stream_f64(a + q * 32, pol) * PE_GATHER(c); }
#undef PE_GATHER
    if (d != 0.0) u[row] += (fr - acc) / d;
}

int pe_launch_sell_gs(pe_ctx *ctx, const DevSELL &S, int s0, int s1, int ext_base, const double *f, double *u,
                      const double *uext, const double *l1)
{
    if (s1 <= s0) return 0;
    if (ctx->rec)
    {
        PeOp o = pe_op(PE_OP_SELL_GS); o.i0 = s0; o.i1 = s1; o.i2 = ext_base;
        o.p[0] = S.soff; o.p[1] = S.J; o.p[2] = S.A; o.p[3] = f; o.p[4] = u; o.p[5] = uext; o.p[6] = l1;
        pe_rec_push(ctx, o, -1.0);
        return 0;
    }
    const int grid = pe_grid_for((int64_t)(s1 - s0) * 32, 256);
    if (uext) k_sell_gs<true><<<grid, 256, 0, ctx->stream>>>(s0, s1, S.soff, S.J, S.A, ext_base, f, u, uext, l1);
    else k_sell_gs<false><<<grid, 256, 0, ctx->stream>>>(s0, s1, S.soff, S.J, S.A, ext_base, f, u, uext, l1);
    PE_LAUNCHED(ctx);
    return 0;
}

// colour-ordered <-> caller numbering
__global__ void k_sell_perm_in(int n, const int *__restrict__ pos, const double *__restrict__ b,
                               const double *__restrict__ x, double *fp, double *up)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pos[i];
    fp[p] = b[i];
    up[p] = x ? x[i] : 0.0;
}
__global__ void k_sell_perm_out(int n, const int *__restrict__ pos, const double *__restrict__ up, double *x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = up[pos[i]];
}
int pe_launch_perm_in(pe_ctx *ctx, int n, const int *pos, const double *b, const double *x, double *fp, double *up)
{
    if (n == 0) return 0;
    if (ctx->rec)
    {
        PeOp o = pe_op(PE_OP_PERM_IN); o.n = n; o.p[0] = pos; o.p[1] = b; o.p[2] = x; o.p[3] = fp; o.p[4] = up;
        pe_rec_push(ctx, o, 0.0);   // renumbering is our overhead, not algorithmic traffic
        return 0;
    }
    k_sell_perm_in<<<pe_grid_for(n, 256), 256, 0, ctx->stream>>>(n, pos, b, x, fp, up);
    PE_LAUNCHED(ctx);
    return 0;
}
int pe_launch_perm_out(pe_ctx *ctx, int n, const int *pos, const double *up, double *x)
{
    if (n == 0) return 0;
    if (ctx->rec)
    {
        PeOp o = pe_op(PE_OP_PERM_OUT); o.n = n; o.p[0] = pos; o.p[1] = up; o.p[2] = x;
        pe_rec_push(ctx, o, 0.0);
        return 0;
    }
    k_sell_perm_out<<<pe_grid_for(n, 256), 256, 0, ctx->stream>>>(n, pos, up, x);
    PE_LAUNCHED(ctx);
    return 0;
}
