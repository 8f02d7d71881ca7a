// prep_kernels.cu -- layout conversion and Fock assembly kernels (all O(n^2) or a
// single pass over the tensor at set-up time; none is on the per-build critical path
// except the O(n^2) ones).
#include "common.cuh"
#include "kernels.cuh"
#include "synth.cuh"

namespace mqcb200 {

// tile id (column-major lower triangle) -> (tr, tc)
__device__ __forceinline__ void tile_coords(int tile, int nt, int &tr, int &tc) {
  // column tc starts at off(tc) = tc*nt - tc*(tc-1)/2; find the largest tc with off(tc) <= tile
  int c = 0, off = 0;
  // closed form via the quadratic, then a correction step for rounding
  const double b = 2.0 * nt + 1.0;
  c = (int)((b - sqrt(b * b - 8.0 * (double)tile)) * 0.5);
  if (c < 0) c = 0;
  if (c > nt - 1) c = nt - 1;
  off = c * nt - (c * (c - 1)) / 2;
  while (off > tile) { --c; off = c * nt - (c * (c - 1)) / 2; }
  while (c + 1 < nt && (c + 1) * nt - ((c + 1) * c) / 2 <= tile) { ++c; off = c * nt - (c * (c - 1)) / 2; }
  tc = c;
  tr = tc + (tile - off);
}

// One CTA (256 threads) per (tile, q): thread -> element (r = tid%16, c = tid/16).
__global__ void __launch_bounds__(256) pack_tensor_kernel(const double *__restrict__ full, int n, int nt,
                                                          long long L, double *__restrict__ packed) {
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int q = blockIdx.y;
  const int r = threadIdx.x & 15, c = threadIdx.x >> 4;
  int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0;
  if (mu < n && nu < n) {
    if (mu < nu) { int t = mu; mu = nu; nu = t; }  // diagonal tile: mirror the lower triangle
    v = full[(size_t)q * n * n + (size_t)mu + (size_t)n * nu];
  }
  packed[(size_t)q * L + (size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(r, c)] = v;
}

__global__ void __launch_bounds__(256) pack_tensor_lower_kernel(const double *__restrict__ lower, int n, int nt,
                                                                long long L, double *__restrict__ packed) {
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int q = blockIdx.y;
  const int r = threadIdx.x & 15, c = threadIdx.x >> 4;
  int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0;
  if (mu < n && nu < n) {
    if (mu < nu) { int t = mu; mu = nu; nu = t; }  // diagonal tile: mirror the lower triangle
    const size_t tri = (size_t)n * (n + 1) / 2;
    v = lower[(size_t)q * tri + (size_t)nu * n - ((size_t)nu * (nu - 1)) / 2 + (size_t)(mu - nu)];
  }
  packed[(size_t)q * L + (size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(r, c)] = v;
}

__global__ void __launch_bounds__(256) synth_tensor_kernel(double *__restrict__ packed, int n, int nt,
                                                           long long L, int q_global_begin, uint64_t seed,
                                                           double inv_width, double scale) {
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int q = blockIdx.y;
  // threads walk the tile in storage order so the 2 KiB store is fully coalesced
  const int o = threadIdx.x;
  const int f = o >> 5, rr = (o & 31) >> 2, cc = o & 3;
  const int r = ((f >> 2) << 3) + rr, c = ((f & 3) << 2) + cc;
  int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0;
  if (mu < n && nu < n) {
    if (mu < nu) { int t = mu; mu = nu; nu = t; }
    v = synth_value(seed, (uint32_t)(q_global_begin + q), (uint32_t)mu, (uint32_t)nu, inv_width, scale);
  }
  packed[(size_t)q * L + (size_t)blockIdx.x * TILE_ELEMS + o] = v;
}

__global__ void __launch_bounds__(256) pack_density_kernel(const double *__restrict__ d, int n, int nt,
                                                           double *__restrict__ w,
                                                           const int *__restrict__ skip_flag, size_t d_stride,
                                                           size_t w_stride) {
  if (skip_flag != nullptr && *skip_flag != 0) return;
  d += (size_t)blockIdx.y * d_stride;               // blockIdx.y: fragment of a batch (strides 0 otherwise)
  w += (size_t)blockIdx.y * w_stride;
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int r = threadIdx.x & 15, c = threadIdx.x >> 4;
  const int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0;
  if (mu < n && nu < n) {
    v = d[(size_t)mu + (size_t)n * nu];
    if (tr != tc) v += d[(size_t)nu + (size_t)n * mu];  // both triangles of an off-diagonal tile
  }
  w[(size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(r, c)] = v;
}

// ctf[((kc*nib + ib)*4 + ks)*32 + g*4 + t] = C[nu = 16kc + 4ks + t][i = 8ib + g]      (B operand of the
//                                                                     half-transform, fragment order)
// cep[(((tr*nib + ib)*2 + rh)*32 + g*4 + t)*2 + e] = C[mu = 16tr + 8rh + g][i = 8ib + 2t + e]
//   (the same matrix in m8n8 ACCUMULATOR order: the half-transform's epilogue dots its X tile with it,
//    one coalesced 16-byte load per accumulator pair)
__global__ void __launch_bounds__(256) pack_coeff_kernel(const double *__restrict__ coeff, int ldc, int n,
                                                         int n_occ, int nib, int nt,
                                                         double *__restrict__ ctf, double *__restrict__ cep,
                                                         size_t coeff_stride, size_t ctf_stride) {
  coeff += (size_t)blockIdx.y * coeff_stride;
  ctf += (size_t)blockIdx.y * ctf_stride;
  const size_t total = (size_t)nt * nib * 128;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = e >> 7;
    const int ib = (int)(blk % nib), kc = (int)(blk / nib);
    {
      const int t = e & 3, g = (e >> 2) & 7, ks = (e >> 5) & 3;
      const int nu = kc * 16 + ks * 4 + t, i = ib * 8 + g;
      ctf[e] = (nu < n && i < n_occ) ? coeff[(size_t)nu + (size_t)ldc * i] : 0.0;
    }
    if (cep != nullptr) {
      const int el = e & 1, t = (e >> 1) & 3, g = (e >> 3) & 7, rh = (e >> 6) & 1;
      const int mu = kc * 16 + rh * 8 + g, i = ib * 8 + 2 * t + el;
      cep[e] = (mu < n && i < n_occ) ? coeff[(size_t)mu + (size_t)ldc * i] : 0.0;
    }
  }
}

// One CTA per 16x16 tile of the LOWER triangle, one thread per element (a >= b) with the
// lanes running along b: both partial buffers are then read in 32-byte sectors that are
// fully used (the K tiles are row-major in b, the packed J rows hold 4 consecutive b per
// fragment row).  Each thread writes its value to (a,b) and to (b,a), so the outputs are
// exactly symmetric.  The sums run over the partials in fixed order.
__global__ void __launch_bounds__(256) finalize_jk_kernel(const double *__restrict__ jpart, int n_jslices,
                                                          long long L, const double *__restrict__ kpart,
                                                          int n_ksplits, int n_ksplits_diag, int n_ktiles,
                                                          int ktile_log2, int n, int nt,
                                                          double k_factor, double *__restrict__ j_out,
                                                          double *__restrict__ k_out,
                                                          const double *__restrict__ h, double jf, double kf,
                                                          double *__restrict__ fock_out, size_t jpart_stride,
                                                          size_t kpart_stride, size_t out_stride, int n_ksplits_edge,
                                                          int edge_panel) {
  if (blockIdx.y != 0) {                            // fragment of a batch: every operand moves on by its stride
    if (jpart) jpart += (size_t)blockIdx.y * jpart_stride;
    if (kpart) kpart += (size_t)blockIdx.y * kpart_stride;
    if (j_out) j_out += (size_t)blockIdx.y * out_stride;
    if (k_out) k_out += (size_t)blockIdx.y * out_stride;
    if (h) h += (size_t)blockIdx.y * out_stride;
    if (fock_out) fock_out += (size_t)blockIdx.y * out_stride;
  }
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int a = tr * 16 + (threadIdx.x >> 4);
  const int b = tc * 16 + (threadIdx.x & 15);
  if (a >= n || b >= n || a < b) return;
  double jv = 0.0, kv = 0.0;
  if (jpart) {
    const size_t off = (size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(a & 15, b & 15);
    // fixed summation order; the loads are independent, so keep many of them in flight
    double s = 0.0;
    int sl = 0;
    for (; sl + 16 <= n_jslices; sl += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = jpart[(size_t)(sl + u) * L + off];
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    for (; sl < n_jslices; ++sl) s += jpart[(size_t)sl * L + off];
    jv = s;
  }
  if (kpart) {
    const int mp = a >> ktile_log2, np = b >> ktile_log2, msk = (1 << ktile_log2) - 1;
    const size_t off = ((size_t)(mp * (mp + 1) / 2 + np) << (2 * ktile_log2)) + ((size_t)(a & msk) << ktile_log2) + (b & msk);
    double s = 0.0;
    const size_t stride = (size_t)n_ktiles << (2 * ktile_log2);
    // diagonal tiles and the tiles of a partly filled last panel row hold fewer partials
    const int n_sp = mp == np ? n_ksplits_diag : (mp == edge_panel ? n_ksplits_edge : n_ksplits);
    int sp = 0;
    for (; sp + 16 <= n_sp; sp += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = kpart[(size_t)(sp + u) * stride + off];
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    for (; sp < n_sp; ++sp) s += kpart[(size_t)sp * stride + off];
    kv = k_factor * s;
  }
  const size_t lo = (size_t)a + (size_t)n * b, up = (size_t)b + (size_t)n * a;
  if (j_out) { j_out[lo] = jv; j_out[up] = jv; }
  if (k_out) { k_out[lo] = kv; k_out[up] = kv; }
  // optional fused assembly F = H + jf*J - kf*K (single-GPU closed-shell builds: saves a launch)
  if (fock_out) {
    const double g2 = jf * jv - kf * kv;
    fock_out[lo] = (h ? h[lo] : 0.0) + g2;
    if (lo != up) fock_out[up] = (h ? h[up] : 0.0) + g2;
  }
}

__global__ void __launch_bounds__(256) assemble_fock_kernel(const double *__restrict__ h,
                                                            const double *__restrict__ j,
                                                            const double *__restrict__ k, double jf, double kf,
                                                            size_t nn, double *__restrict__ f) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    double v = h ? h[e] : 0.0;
    if (j) v += jf * j[e];
    if (k) v -= kf * k[e];
    f[e] = v;
  }
}

__global__ void __launch_bounds__(256) combine_g_kernel(const double *__restrict__ j, const double *__restrict__ ka,
                                                        const double *__restrict__ kb, double ca, double cb,
                                                        size_t nn, double *__restrict__ g) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    double v = j ? j[e] : 0.0;
    if (ka) v += ca * ka[e];
    if (kb) v += cb * kb[e];
    g[e] = v;
  }
}

// E = 1/2 sum D (H + F).  Each block reduces a fixed slice of the matrix to one partial; the
// block that finishes last adds the partials in index order, so the result does not depend
// on scheduling.  scratch: [128 partial slots | 1 counter at a FIXED slot (left at 0)].
__global__ void __launch_bounds__(256) energy_kernel(const double *__restrict__ d, const double *__restrict__ h,
                                                     const double *__restrict__ f, size_t nn,
                                                     double *__restrict__ scratch, double *__restrict__ out,
                                                     size_t mat_stride, size_t out_stride) {
  d += (size_t)blockIdx.y * mat_stride;             // fragment of a batch: own matrices, own scratch (130 doubles), own result
  h += (size_t)blockIdx.y * mat_stride;
  f += (size_t)blockIdx.y * mat_stride;
  scratch += (size_t)blockIdx.y * 130;
  out += (size_t)blockIdx.y * out_stride;
  __shared__ double red[8];
  __shared__ bool last;
  double s = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
    s += d[e] * (h[e] + f[e]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    scratch[blockIdx.x] = t;
    __threadfence();
    unsigned int *counter = reinterpret_cast<unsigned int *>(scratch + 128);
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += reinterpret_cast<volatile double *>(scratch)[b];
    out[0] = 0.5 * t;
    *reinterpret_cast<unsigned int *>(scratch + 128) = 0u;   // ready for the next build
  }
}

// ------------------------------- launchers ----------------------------------
void launch_pack_tensor(const double *d_full, int n, int q_count, double *d_packed, cudaStream_t s) {
  const int nt = num_tiles(n);
  const long long L = packed_row_len(n);
  for (int q0 = 0; q0 < q_count; q0 += 65535) {
    const int qc = q_count - q0 < 65535 ? q_count - q0 : 65535;
    dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)qc);
    pack_tensor_kernel<<<grid, 256, 0, s>>>(d_full + (size_t)q0 * n * n, n, nt, L, d_packed + (size_t)q0 * L);
  }
}

void launch_pack_tensor_lower(const double *d_lower, int n, int q_count, double *d_packed, cudaStream_t s) {
  const int nt = num_tiles(n);
  const long long L = packed_row_len(n);
  const size_t tri = (size_t)n * (n + 1) / 2;
  for (int q0 = 0; q0 < q_count; q0 += 65535) {
    const int qc = q_count - q0 < 65535 ? q_count - q0 : 65535;
    dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)qc);
    pack_tensor_lower_kernel<<<grid, 256, 0, s>>>(d_lower + (size_t)q0 * tri, n, nt, L, d_packed + (size_t)q0 * L);
  }
}

void launch_synth_tensor(double *d_packed, int n, int q_global_begin, int q_count, uint64_t seed,
                         double scale, cudaStream_t s) {
  const int nt = num_tiles(n);
  const long long L = packed_row_len(n);
  const double inv_width = 8.0 / (double)n;
  for (int q0 = 0; q0 < q_count; q0 += 65535) {
    const int qc = q_count - q0 < 65535 ? q_count - q0 : 65535;
    dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)qc);
    synth_tensor_kernel<<<grid, 256, 0, s>>>(d_packed + (size_t)q0 * L, n, nt, L, q_global_begin + q0, seed,
                                             inv_width, scale);
  }
}

void launch_pack_density(const double *d_density, int n, double *d_w, const int *d_skip_flag, cudaStream_t s, int batch,
                         size_t d_stride, size_t w_stride) {
  const int nt = num_tiles(n);
  dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)batch);
  pack_density_kernel<<<grid, 256, 0, s>>>(d_density, n, nt, d_w, d_skip_flag, d_stride, w_stride);
}

// Density packing (w[L], as pack_density_kernel) fused with the on-device decision whether the
// density is the orbitals' own, D == f (Ca Ca^T + Cb Cb^T): one CTA per lower-triangular tile,
// one thread per element, BOTH triangles of D compared (an asymmetric D is not an orbital
// product).  Block maxima go through integer atomicMax (non-negative doubles order as integers;
// a NaN/Inf anywhere is mapped to +Inf so it can never be dropped by a max), the block that
// finishes last turns them into the flag and re-zeroes the scratch for the next build.
// scratch: [0] max residual, [1] max |D| (u64 each), [2] u32 arrival counter.
__global__ void __launch_bounds__(256) density_prep_kernel(const double *__restrict__ d, int n, int nt,
                                                           double *__restrict__ w,
                                                           const double *__restrict__ ca, int lda, int na,
                                                           const double *__restrict__ cb, int ldb, int nb,
                                                           double f, double tol_rel,
                                                           unsigned long long *__restrict__ scratch,
                                                           int *__restrict__ flag) {
  int tr, tc;
  tile_coords(blockIdx.x, nt, tr, tc);
  const int r = threadIdx.x & 15, c = threadIdx.x >> 4;
  const int mu = tr * TILE + r, nu = tc * TILE + c;
  double v = 0.0, resid = 0.0, dabs = 0.0;
  if (mu < n && nu < n) {
    const double lo = d[(size_t)mu + (size_t)n * nu];
    const double up = tr != tc ? d[(size_t)nu + (size_t)n * mu] : lo;
    v = tr != tc ? lo + up : lo;
    double s = 0.0;
    for (int i = 0; i < na; ++i) s = fma(ca[(size_t)mu + (size_t)lda * i], ca[(size_t)nu + (size_t)lda * i], s);
    for (int i = 0; i < nb; ++i) s = fma(cb[(size_t)mu + (size_t)ldb * i], cb[(size_t)nu + (size_t)ldb * i], s);
    resid = fmax(fabs(lo - f * s), fabs(up - f * s));
    dabs = fmax(fabs(lo), fabs(up));
    if (!isfinite(lo) || !isfinite(up) || !isfinite(s)) resid = __longlong_as_double(0x7ff0000000000000ll);
  }
  w[(size_t)blockIdx.x * TILE_ELEMS + in_tile_offset(r, c)] = v;
  for (int o = 16; o > 0; o >>= 1) {
    resid = fmax(resid, __shfl_xor_sync(0xffffffffu, resid, o));
    dabs = fmax(dabs, __shfl_xor_sync(0xffffffffu, dabs, o));
  }
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) {
    atomicMax(scratch, (unsigned long long)__double_as_longlong(resid));
    atomicMax(scratch + 1, (unsigned long long)__double_as_longlong(dabs));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int *>(scratch + 2), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const double rmax = __longlong_as_double((long long)atomicExch(scratch, 0ull));
    const double dmax = __longlong_as_double((long long)atomicExch(scratch + 1, 0ull));
    *flag = (rmax <= tol_rel * fmax(1.0, dmax)) ? 1 : 0;
    *reinterpret_cast<unsigned int *>(scratch + 2) = 0u;
  }
}

void launch_density_prep(const double *d_density, int n, double *d_w, const double *d_ca, int lda, int na,
                         const double *d_cb, int ldb, int nb, double f, unsigned long long *d_scratch, int *d_flag,
                         cudaStream_t s) {
  const int nt = num_tiles(n);
  const int occ = na + nb > 16 ? na + nb : 16;
  const double tol_rel = 4.0 * 1.1102230246251565e-16 * (double)occ;   // a few ulp per term of the orbital sum
  density_prep_kernel<<<(unsigned)num_lower_tiles(nt), 256, 0, s>>>(d_density, n, nt, d_w, d_ca, lda, na, d_cb, ldb, nb, f,
                                                                    tol_rel, d_scratch, d_flag);
}

__global__ void __launch_bounds__(256) stack_factors_kernel(const double *__restrict__ x, int ldx,
                                                            const double *__restrict__ c, int ldc, int n, int n_occ,
                                                            int o16, double *__restrict__ s) {
  const size_t total = (size_t)n * 2 * o16;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int mu = (int)(e % n), col = (int)(e / n);
    const int i = col < o16 ? col : col - o16;
    double v = 0.0;
    if (i < n_occ) v = col < o16 ? x[(size_t)mu + (size_t)ldx * i] : c[(size_t)mu + (size_t)ldc * i];
    s[e] = v;
  }
}

void launch_stack_factors(const double *d_x, int ldx, const double *d_c, int ldc, int n, int n_occ, double *d_s,
                          cudaStream_t s) {
  const int o16 = (n_occ + 15) / 16 * 16;
  const size_t total = (size_t)n * 2 * o16;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  stack_factors_kernel<<<blocks, 256, 0, s>>>(d_x, ldx, d_c, ldc, n, n_occ, o16, d_s);
}

void launch_pack_coeff(const double *d_coeff, int ldc, int n, int n_occ, int nib, double *d_ctf, double *d_cep,
                       cudaStream_t s, int batch, size_t coeff_stride, size_t ctf_stride) {
  const int nt = num_tiles(n);
  const size_t total = (size_t)nt * nib * 128;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid(blocks, (unsigned)batch);
  pack_coeff_kernel<<<grid, 256, 0, s>>>(d_coeff, ldc, n, n_occ, nib, nt, d_ctf, d_cep, coeff_stride, ctf_stride);
}

void launch_finalize_jk(const double *d_jpart, int n_jslices, const double *d_kpart, int n_ksplits, int ktile,
                        int n, double k_factor, double *d_j, double *d_k, cudaStream_t s, const double *d_h,
                        double jf, double kf, double *d_fock, int n_ksplits_diag, int batch, size_t jpart_stride,
                        size_t kpart_stride, size_t out_stride, int n_ksplits_edge) {
  const int nt = num_tiles(n);
  const int ktile_log2 = ktile == 128 ? 7 : 6;
  const int np = (n + ktile - 1) / ktile;
  dim3 grid((unsigned)num_lower_tiles(nt), (unsigned)batch);
  finalize_jk_kernel<<<grid, 256, 0, s>>>(d_jpart, n_jslices, packed_row_len(n), d_kpart, n_ksplits,
                                          n_ksplits_diag < 0 ? n_ksplits : n_ksplits_diag,
                                          np * (np + 1) / 2, ktile_log2, n, nt, k_factor, d_j, d_k, d_h, jf, kf,
                                          d_fock, jpart_stride, kpart_stride, out_stride,
                                          n_ksplits_edge < 0 ? n_ksplits : n_ksplits_edge,
                                          (n_ksplits_edge < 0 || n_ksplits_edge == n_ksplits) ? -1 : np - 1);
}

void launch_assemble_fock(const double *d_h, const double *d_j, const double *d_k, double jf, double kf, int n,
                          double *d_fock, cudaStream_t s) {
  const size_t nn = (size_t)n * n;
  unsigned blocks = (unsigned)((nn + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  assemble_fock_kernel<<<blocks, 256, 0, s>>>(d_h, d_j, d_k, jf, kf, nn, d_fock);
}

void launch_combine_g(const double *d_j, const double *d_ka, const double *d_kb, double ka, double kb, int n,
                      double *d_g, cudaStream_t s) {
  const size_t nn = (size_t)n * n;
  unsigned blocks = (unsigned)((nn + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  combine_g_kernel<<<blocks, 256, 0, s>>>(d_j, d_ka, d_kb, ka, kb, nn, d_g);
}

void launch_energy(const double *d_density, const double *d_h, const double *d_fock, int n, double *d_scratch,
                   double *d_out, cudaStream_t s, int batch, size_t mat_stride, size_t out_stride) {
  const size_t nn = (size_t)n * n;
  unsigned blocks = (unsigned)((nn + 4095) / 4096);
  if (blocks > 128) blocks = 128;
  if (blocks < 1) blocks = 1;
  dim3 grid(blocks, (unsigned)batch);
  energy_kernel<<<grid, 256, 0, s>>>(d_density, d_h, d_fock, nn, d_scratch, d_out, mat_stride, out_stride);
}

}  // namespace mqcb200
