// Host-side invariants of the launch planners and of the packed-layout arithmetic, over a
// grid of shapes far wider than the GPU parity suite.  Built and run on the CPU by
// tests/test_native_planners.py (links libmqcb200.so; launches nothing).
#include <sys/mman.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <utility>
#include <vector>

#include "../../metalquicha_b200/csrc/common.cuh"
#include "../../metalquicha_b200/csrc/kernels.cuh"

using namespace mqcb200;

static int failures = 0;
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      ++failures;                                         \
      if (failures < 20) { printf("FAIL %s: ", #cond); printf(__VA_ARGS__); printf("\n"); } \
    }                                                     \
  } while (0)

static void check_layout() {
  std::set<int> seen;
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 16; ++c) {
      const int o = in_tile_offset(r, c);
      CHECK(o >= 0 && o < 256, "offset %d", o);
      seen.insert(o);
    }
  CHECK(seen.size() == 256, "in_tile_offset is not a permutation");
  for (int nt = 1; nt <= 100; ++nt) {
    std::set<int> ids;
    for (int tc = 0; tc < nt; ++tc)
      for (int tr = tc; tr < nt; ++tr) ids.insert(tile_index(tr, tc, nt));
    CHECK((long long)ids.size() == num_lower_tiles(nt) && *ids.begin() == 0 &&
              *ids.rbegin() == (int)num_lower_tiles(nt) - 1,
          "tile_index not a bijection at nt=%d", nt);
  }
  CHECK(packed_row_len(688) == 946LL * 256 && packed_row_len(1450) == 4186LL * 256, "row lengths");
}

// the unit -> (tile, split) decode k_accumulate_kernel uses (common.cuh: one function for both sides)
static void decode_unit(int u, const KPlan &p, int &tile, int &split, int &mp, int &np) {
  k_unit_decode(u, p.n_ktiles, p.n_panels, p.n_splits, p.n_splits_diag, p.n_splits_edge, mp, np, split);
  tile = mp * (mp + 1) / 2 + np;
}
static int splits_of(const KPlan &p, int mp, int np) {      // what k_accumulate_kernel and finalize_jk_kernel use
  return mp == np ? p.n_splits_diag : (mp == p.n_panels - 1 ? p.n_splits_edge : p.n_splits);
}

static void check_plan_k() {
  const int ns[] = {1, 8, 16, 17, 24, 63, 64, 65, 72, 81, 100, 127, 128, 129, 200, 257, 688, 1000, 1450, 2047, 4096};
  const int os[] = {1, 5, 8, 15, 16, 17, 64, 65, 80, 96, 113, 128, 129, 144, 241, 256, 257, 400};
  const int qs[] = {1, 3, 7, 64, 225, 340, 1800, 6800};
  const size_t wss[] = {(size_t)1 << 20, (size_t)64 << 20, (size_t)4 << 30};
  for (int n : ns)
    for (int o : os) {
      if (o > n) continue;
      for (int q : qs)
        for (size_t ws : wss) {
          const KPlan p = plan_k(n, o, q, ws, 148);
          const int nt = num_tiles(n);
          CHECK(p.nb >= 1 && p.nb <= 8, "nb=%d (n=%d o=%d)", p.nb, n, o);
          CHECK(p.nib == p.n_ntiles * 2 * p.nb && p.nib * 8 >= o, "nib=%d ntiles=%d nb=%d o=%d", p.nib, p.n_ntiles, p.nb, o);
          CHECK(p.nkc * 2 == p.nib, "nkc=%d nib=%d", p.nkc, p.nib);
          CHECK(p.nib * 8 - o < 16 * p.n_ntiles + 16, "padding of the occupied range too large: nib=%d o=%d", p.nib, o);
          CHECK(p.nmb == 2 * nt, "nmb");
          CHECK(p.ks_last >= 0 && p.ks_last <= 4, "ks_last=%d", p.ks_last);
          // the k-subs the accumulation kernel skips must all be padding
          CHECK(16 * (p.nkc - 1) + 4 * p.ks_last >= o || p.ks_last == 4, "ks_last=%d skips valid columns (o=%d nkc=%d)", p.ks_last, o, p.nkc);
          CHECK(p.ktile == 64 || p.ktile == 128, "ktile");
          CHECK(p.n_panels == (n + p.ktile - 1) / p.ktile && p.n_ktiles == p.n_panels * (p.n_panels + 1) / 2, "tiles");
          CHECK(p.n_splits >= 1 && p.n_splits <= q && p.n_splits_diag >= 1 && p.n_splits_diag <= p.n_splits,
                "splits %d/%d q=%d", p.n_splits, p.n_splits_diag, q);
          CHECK(p.q_chunk >= 1 && p.q_chunk <= q, "q_chunk=%d", p.q_chunk);
          CHECK(p.q_chunk == 1 || p.x_elems_per_q * (size_t)p.q_chunk * 8 <= ws, "workspace overrun");
          CHECK(p.x_elems_per_q == (size_t)p.nkc * p.nmb * 128, "x size");
          CHECK(p.kpart_elems == (size_t)p.n_splits * p.n_ktiles * p.ktile * p.ktile, "kpart size");
          CHECK(p.kpart_elems * 8 <= ((size_t)256 << 20) || p.n_splits == 1, "partial buffer cap");
          CHECK(p.gamma_stride == nt * p.n_ntiles * 2, "gamma stride");
          // a trimmed last n8-block holds only padding, and the accumulation stops short of its columns
          CHECK(!p.trim_last || (p.nb > 1 && p.nib * 8 - o >= 8 && 16 * (p.nkc - 1) + 4 * p.ks_last <= p.nib * 8 - 8),
                "trim_last reaches live columns (o=%d nib=%d nkc=%d ks_last=%d)", o, p.nib, p.nkc, p.ks_last);
          CHECK(p.trim_last || p.nb == 1 || p.nib * 8 - o < 8, "a whole padding block is left untrimmed (o=%d nib=%d)", o, p.nib);
          // the last panel row gets its own split count only when it is partly filled, in proportion to its live row blocks
          {
            const int lb = p.nmb - 8 * (p.n_panels - 1);
            CHECK(p.n_splits_edge >= 1 && p.n_splits_edge <= p.n_splits, "edge splits %d of %d", p.n_splits_edge, p.n_splits);
            if (p.ktile != 64 || lb >= 8 || p.n_panels < 2) CHECK(p.n_splits_edge == p.n_splits, "edge class without an edge (n=%d)", n);
            else CHECK(p.n_splits_edge == p.n_splits || p.n_splits_edge == (lb * p.n_splits + 4) / 8 || p.n_splits_edge == 1,
                       "edge splits %d for lb=%d s=%d", p.n_splits_edge, lb, p.n_splits);
          }
          // every (tile, split < its split count) is produced by exactly one unit
          const long long units = k_unit_count(p.n_ktiles, p.n_panels, p.n_splits, p.n_splits_diag, p.n_splits_edge);
          if (units <= 200000) {
            std::set<std::pair<int, int>> got;
            bool ok = true;
            for (int u = 0; u < (int)units; ++u) {
              int tile, split, mp, np;
              decode_unit(u, p, tile, split, mp, np);
              ok = ok && np >= 0 && np <= mp && mp < p.n_panels && split >= 0 && split < splits_of(p, mp, np);
              ok = ok && got.insert({tile, split}).second;
            }
            long long expect = 0;
            for (int mp = 0; mp < p.n_panels; ++mp)
              for (int np = 0; np <= mp; ++np) expect += splits_of(p, mp, np);
            CHECK(ok && (long long)got.size() == expect && units == expect, "unit decode (n=%d o=%d q=%d): %lld units, %d distinct, %lld expected",
                  n, o, q, units, (int)got.size(), expect);
          }
        }
    }
}

// rank-2 (SYR2K-form) plans: the stacked operand has 2*ceil16(o) columns, X-chunk k pairs with chunk
// k + pair_off of the C half, every tile (diagonal ones too) takes the full two-panel path
static void check_plan_rank2() {
  for (int n : {17, 58, 72, 130, 688, 1450})
    for (int o : {1, 9, 15, 16, 17, 33, 80, 81, 129, 241}) {
      if (o > n) continue;
      for (int q : {1, 12, 70, 340, 1800}) {
        const KPlan p = plan_k(n, /*ignored*/ 0, q, (size_t)4 << 30, 148, o);
        const int o16 = (o + 15) / 16 * 16;
        CHECK(p.pair_off == o16 / 16 && p.npairs == o16 / 16, "pairs (o=%d): off=%d n=%d", o, p.pair_off, p.npairs);
        CHECK(p.nib * 8 >= 2 * o16 && p.nkc * 2 == p.nib, "stacked width nib=%d o16=%d", p.nib, o16);
        CHECK(p.pair_off + p.npairs <= p.nkc, "the C half must lie inside the half-transformed chunks");
        CHECK(p.ks_last >= 1 && p.ks_last <= 4 && 16 * (p.npairs - 1) + 4 * p.ks_last >= o, "ks_last=%d o=%d", p.ks_last, o);
        CHECK(!p.trim_last || 16 * (p.pair_off + p.npairs) <= p.nib * 8 - 8, "rank-2 trim_last reaches live columns (o=%d nib=%d)", o, p.nib);
        CHECK(p.n_splits_diag == p.n_splits, "rank-2 diagonal tiles are full tiles: %d vs %d", p.n_splits_diag, p.n_splits);
        CHECK(p.n_splits_edge == p.n_splits, "rank-2 plans keep the two-class schedule: %d vs %d", p.n_splits_edge, p.n_splits);
        CHECK(p.kpart_elems == (size_t)p.n_splits * p.n_ktiles * p.ktile * p.ktile, "kpart size");
      }
    }
}

static void check_plan_j_and_fragment() {
  for (int n : {1, 16, 24, 72, 80, 81, 688, 1450})
    for (int q : {1, 7, 64, 340, 1800, 6800}) {
      const JPlan p = plan_j(n, q);
      const long long L = packed_row_len(n);
      CHECK(p.n_seg >= 1 && p.n_slices >= 1 && p.n_slices <= (q > 0 ? q : 1), "j plan n=%d q=%d", n, q);
      CHECK(p.gamma_partial_elems == (size_t)p.n_seg * q && p.j_partial_elems == (size_t)p.n_slices * L, "j buffers");
    }
  for (int n = 1; n <= 80; ++n)
    for (int o : {1, 5, 8, 9, 15, 33, 64}) {
      if (o > n) continue;
      CHECK(fragment_path_applies(n, o), "fragment path refused n=%d o=%d", n, o);
      for (int q : {1, 5, 113, 340, 5000}) {
        const FragPlan p = plan_fragment(n, o, q, 148);
        CHECK(p.nt == num_tiles(n) && p.nt <= 5 && p.nib == (o + 7) / 8 && p.nib <= 8, "fragment shape");
        CHECK(p.grid >= 1 && p.grid <= 148 && p.grid <= q, "fragment grid %d (q=%d)", p.grid, q);
        CHECK(p.smem_bytes <= 200 * 1024, "fragment smem %zu", p.smem_bytes);
        CHECK(p.L == (int)packed_row_len(n) && p.jpart_elems == (size_t)p.grid * p.L, "fragment buffers");
        CHECK(p.kpart_elems == (size_t)p.grid * p.n_ktiles * 4096, "fragment K buffer");
      }
    }
  CHECK(!fragment_path_applies(81, 5) && !fragment_path_applies(72, 65), "fragment path limits");
}

// the half-transform's last round: whole rounds stay whole, the pieces of the rest fit ONE round, and the
// (unit, warp column, slot) pieces the kernel decodes cover every unit of the rest exactly once
static void check_half_tail() {
  for (int ctas : {1, 2, 100, 132, 148})
    for (long long units = 0; units <= 5000; units += (units < 700 ? 1 : 37)) {
      const HalfTail t = plan_half_tail(units, ctas);
      CHECK(t.split == 1 || t.split == 2 || t.split == 4, "split=%d", t.split);
      CHECK(t.n_full >= 0 && t.n_full <= units && t.n_full % ctas == 0 || t.split == 1, "n_full=%lld units=%lld ctas=%d", t.n_full, units, ctas);
      if (t.split == 1) { CHECK(t.n_full == units, "no split but n_full=%lld != %lld", t.n_full, units); continue; }
      const long long rest = units - t.n_full;
      CHECK(rest > 0 && rest < ctas && rest * t.split <= ctas, "pieces %lld x %d do not fit %d CTAs", rest, t.split, ctas);
      CHECK(t.split == 4 || rest * 4 > ctas, "a finer split would have fitted (rest=%lld ctas=%d)", rest, ctas);
      std::set<long long> seen;
      for (long long work = t.n_full; work < t.n_full + rest * t.split; ++work) {       // the kernel's decode
        const long long r = work - t.n_full, unit = t.n_full + r / t.split;
        const int p = (int)(r % t.split), wn = p & 1, sl = t.split == 4 ? (p >> 1) : -1;
        for (int s = 0; s < 2; ++s)
          if (sl < 0 || sl == s) seen.insert((unit * 2 + wn) * 2 + s);
      }
      CHECK((long long)seen.size() == rest * 4 && *seen.begin() == t.n_full * 4, "tail pieces do not tile the rest (units=%lld ctas=%d)", units, ctas);
    }
}

// gather_lower (the host side of set_tensor) copies columns in fixed 64-byte pieces that run over into the
// next column's place.  Source and destination are laid out so that they END on a page the process may not
// touch: a read past the caller's array or a write past the triangles faults instead of passing silently.
struct Guarded {
  char *base = nullptr;
  size_t map_bytes = 0;
  double *p = nullptr;
  explicit Guarded(size_t n_doubles) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t bytes = n_doubles * sizeof(double);
    const size_t data_pages = (bytes + page - 1) / page + 1;
    map_bytes = (data_pages + 1) * page;
    base = static_cast<char *>(mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (base == MAP_FAILED) { perror("mmap"); exit(2); }
    mprotect(base + data_pages * page, page, PROT_NONE);
    p = reinterpret_cast<double *>(base + data_pages * page - bytes);
  }
  ~Guarded() { munmap(base, map_bytes); }
};

static void check_gather_lower() {
  const int ns[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 16, 17, 24, 31, 72};
  const size_t qs[] = {1, 2, 3, 5, 17};
  for (int n : ns)
    for (size_t q : qs) {
      const size_t nn = (size_t)n * n, tri = (size_t)n * (n + 1) / 2;
      Guarded src(q * nn), dst(q * tri);
      for (size_t i = 0; i < q * nn; ++i) src.p[i] = (double)(i + 1);
      gather_lower(src.p, n, q, nn, dst.p);
      bool ok = true;
      for (size_t s = 0; s < q && ok; ++s) {
        size_t off = 0;
        for (int nu = 0; nu < n && ok; ++nu)
          for (int mu = nu; mu < n; ++mu, ++off) ok = ok && dst.p[s * tri + off] == src.p[s * nn + (size_t)nu * n + mu];
      }
      CHECK(ok, "gather_lower n=%d q=%zu", n, q);
    }
  {  // large enough (>= 16 MiB of triangles) for the threaded split of the slab range
    const int n = 72;
    const size_t q = 820, nn = (size_t)n * n, tri = (size_t)n * (n + 1) / 2;
    Guarded src(q * nn), dst(q * tri);
    for (size_t i = 0; i < q * nn; ++i) src.p[i] = (double)(i % 1000003);
    gather_lower(src.p, n, q, nn, dst.p);
    bool ok = true;
    for (size_t s = 0; s < q && ok; ++s) {
      size_t off = 0;
      for (int nu = 0; nu < n && ok; ++nu)
        for (int mu = nu; mu < n; ++mu, ++off) ok = ok && dst.p[s * tri + off] == src.p[s * nn + (size_t)nu * n + mu];
    }
    CHECK(ok, "gather_lower threaded n=%d q=%zu", n, q);
  }
}

// ---- diis_solve_host against the reference's own DIIS test -------------------------------------------
// test/test_mqc_diis.f90 of the reference compares diis_state_t, step by step, with the shift-the-history,
// rebuild-B algorithm it replaced (reference_extrapolate, :256-312) on vectors from its own LCG filler
// (:314-332), to 1e-12 * max(1, |F|) (:10).  The same case, same sizes and seeds (:157-223), is run here
// on the engine's host-side solve; the ring bookkeeping around it is what scf_general does.
static void fill_pseudorandom(std::vector<double> &a, int rows, int cols, int seed) {
  a.assign((size_t)rows * cols, 0.0);
  long long state = seed;
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) {
      state = (1103515245LL * state + 12345LL) % 2147483648LL;
      a[(size_t)j * rows + i] = (double)state / 2147483648.0 - 0.5;
    }
}

static bool reference_extrapolate(const std::vector<double> &hist_f, const std::vector<double> &hist_e, int nf, int ne,
                                  int n_stored, std::vector<double> &fock) {
  fock.assign(nf, 0.0);
  if (n_stored < 2) return false;
  const int n = n_stored + 1;
  std::vector<double> aug((size_t)n * (n + 1), 0.0);
  auto A = [&](int r, int c) -> double & { return aug[(size_t)r * (n + 1) + c]; };
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A(i, j) = -1.0;
  A(n - 1, n - 1) = 0.0;
  for (int i = 0; i < n_stored; ++i)
    for (int j = 0; j < n_stored; ++j) {
      double s = 0.0;
      for (int k = 0; k < ne; ++k) s += hist_e[(size_t)i * ne + k] * hist_e[(size_t)j * ne + k];
      A(i, j) = s;
    }
  A(n - 1, n) = -1.0;
  for (int i = 0; i < n; ++i) {
    int piv = i;
    for (int j = i + 1; j < n; ++j)
      if (std::fabs(A(j, i)) > std::fabs(A(piv, i))) piv = j;
    if (piv != i)
      for (int c = 0; c <= n; ++c) std::swap(A(i, c), A(piv, c));
    const double pivot = A(i, i);
    if (std::fabs(pivot) < 1.0e-14) return false;
    for (int j = i + 1; j < n; ++j) {
      const double factor = A(j, i) / pivot;
      for (int c = i; c <= n; ++c) A(j, c) -= factor * A(i, c);
    }
  }
  std::vector<double> coef(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double s = 0.0;
    for (int c = i + 1; c < n; ++c) s += A(i, c) * coef[c];
    coef[i] = (A(i, n) - s) / A(i, i);
  }
  for (int i = 0; i < n_stored; ++i)
    for (int k = 0; k < nf; ++k) fock[k] += coef[i] * hist_f[(size_t)i * nf + k];
  return true;
}

static void check_diis_solve() {
  struct Case { int nf, ne, nmax, npush, seed_f, seed_e; };
  const Case cases[] = {{12, 12, 4, 9, 11, 29},      // test_vs_reference
                        {6, 6, 3, 10, 17, 23},       // test_ring_evicts' data
                        {7, 7, 3, 8, 13, 5},         // test_overlap_cache's data
                        {20, 20, 8, 20, 3, 31}};     // the ring size the SCF uses
  for (const Case &c : cases) {
    std::vector<double> f_in, e_in;
    fill_pseudorandom(f_in, c.nf, c.npush, c.seed_f);
    fill_pseudorandom(e_in, c.ne, c.npush, c.seed_e);
    std::vector<double> ring_f((size_t)c.nmax * c.nf), ring_e((size_t)c.nmax * c.ne);     // by slot
    std::vector<double> hist_f((size_t)c.nmax * c.nf), hist_e((size_t)c.nmax * c.ne);     // by age (the reference's)
    double overlap[64] = {0.0};
    int n_stored = 0, newest = 0, n_hist = 0;
    for (int step = 0; step < c.npush; ++step) {
      // the ring push of scf_general (diis_push, mqc_diis.f90:94-119): overlaps of the newest vector only
      newest = newest % c.nmax + 1;
      if (n_stored < c.nmax) ++n_stored;
      const int slot = newest - 1;
      for (int k = 0; k < c.nf; ++k) ring_f[(size_t)slot * c.nf + k] = f_in[(size_t)step * c.nf + k];
      for (int k = 0; k < c.ne; ++k) ring_e[(size_t)slot * c.ne + k] = e_in[(size_t)step * c.ne + k];
      for (int age = 0; age < n_stored; ++age) {
        const int other = ((newest - n_stored + age) % c.nmax + c.nmax) % c.nmax;
        double s = 0.0;
        for (int k = 0; k < c.ne; ++k) s += ring_e[(size_t)slot * c.ne + k] * ring_e[(size_t)other * c.ne + k];
        overlap[slot * 8 + other] = s;
        overlap[other * 8 + slot] = s;
      }
      // the reference's history: shift down when full, append at the end
      if (n_hist < c.nmax) ++n_hist;
      else {
        for (int i = 0; i + 1 < c.nmax; ++i) {
          for (int k = 0; k < c.nf; ++k) hist_f[(size_t)i * c.nf + k] = hist_f[(size_t)(i + 1) * c.nf + k];
          for (int k = 0; k < c.ne; ++k) hist_e[(size_t)i * c.ne + k] = hist_e[(size_t)(i + 1) * c.ne + k];
        }
      }
      for (int k = 0; k < c.nf; ++k) hist_f[(size_t)(n_hist - 1) * c.nf + k] = f_in[(size_t)step * c.nf + k];
      for (int k = 0; k < c.ne; ++k) hist_e[(size_t)(n_hist - 1) * c.ne + k] = e_in[(size_t)step * c.ne + k];

      double coef[8];
      int slots[8];
      const bool ok = diis_solve_host(overlap, newest, n_stored, c.nmax, coef, slots);
      std::vector<double> ref;
      const bool ok_ref = reference_extrapolate(hist_f, hist_e, c.nf, c.ne, n_hist, ref);
      CHECK(ok == ok_ref, "DIIS solvability disagrees at step %d (nmax=%d)", step + 1, c.nmax);
      if (!ok || !ok_ref) continue;
      std::vector<double> fock(c.nf, 0.0);
      double sum = 0.0;
      for (int i = 0; i < n_stored; ++i) {
        sum += coef[i];
        for (int k = 0; k < c.nf; ++k) fock[k] += coef[i] * ring_f[(size_t)slots[i] * c.nf + k];
        // oldest first: age i of the ring is entry i of the shifted history
        for (int k = 0; k < c.ne; ++k)
          CHECK(ring_e[(size_t)slots[i] * c.ne + k] == hist_e[(size_t)i * c.ne + k], "slot order at step %d", step + 1);
      }
      double dev = 0.0, scale = 1.0;
      for (int k = 0; k < c.nf; ++k) {
        dev = std::fmax(dev, std::fabs(fock[k] - ref[k]));
        scale = std::fmax(scale, std::fabs(ref[k]));
      }
      CHECK(dev <= 1.0e-12 * scale, "extrapolated Fock differs from the reference algorithm at step %d by %.3e (nmax=%d)", step + 1, dev, c.nmax);
      CHECK(std::fabs(sum - 1.0) <= 1.0e-10, "DIIS weights sum to %.15f", sum);
    }
  }
  {  // below two vectors nothing is extrapolated (test_below_two)
    double overlap[64] = {0.0}, coef[8];
    int slots[8];
    overlap[0] = 1.0;
    CHECK(!diis_solve_host(overlap, 1, 0, 4, coef, slots) && !diis_solve_host(overlap, 1, 1, 4, coef, slots), "extrapolated below two vectors");
  }
}

int main() {
  check_gather_lower();
  check_diis_solve();
  check_layout();
  check_half_tail();
  check_plan_k();
  check_plan_rank2();
  check_plan_j_and_fragment();
  if (failures) { printf("%d planner/layout checks FAILED\n", failures); return 1; }
  printf("planner and layout invariants hold\n");
  return 0;
}
