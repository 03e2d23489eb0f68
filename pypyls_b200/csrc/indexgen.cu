// On-device resample tables.  Replaces gen_permsamp / gen_bootsamp
// (pyls/base.py:10-79, 82-159) with a counter-based generator: the stream of
// resample `id` is Philox4x32-10 keyed by the user's seed with counter
// (block, id, attempt, kind), so any rank can produce any column and the table
// does not depend on how resamples are sharded.  Validity rules are the
// reference's:
//   permutation  every subject keeps its conditions together, in an
//                independently shuffled order (utils.permute_cols,
//                pyls/utils.py:200-224); subjects are permuted across groups and
//                a draw in which some group keeps its own subject set is
//                rejected (base.py:59-62); a column equal to an earlier one is
//                re-drawn (base.py:67-69)
//   bootstrap    sorted sampling with replacement inside each group, at least
//                ceil(min(groups)/2) distinct subjects per group (base.py:111,
//                134-139), conditions follow their subject (:142-143); a column
//                whose rows [a,b) equal an earlier column's for some group
//                bounds (a,b) is re-drawn (:145-149, bounds in SUBJECT units
//                as in the reference)
//   at most 500 draws per column (base.py:51,125); columns that hit the cap are
//   kept and counted.
// One thread generates one column; duplicate detection compares 64-bit hashes
// first and full columns on a hash match.
#include "common.cuh"

namespace plsb {
namespace {

constexpr int MAX_TRIES = 500;

struct Philox {
  uint32_t k0, k1, c1, c2, c3, n;
  uint32_t out[4];
  int have;
  __device__ Philox(uint64_t seed, uint32_t id, uint32_t attempt, uint32_t kind)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c1(id), c2(attempt), c3(kind), n(0),
        have(0) {}
  __device__ void refill() {
    uint32_t a = n++, b = c1, c = c2, d = c3, key0 = k0, key1 = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, a), lo0 = 0xD2511F53u * a;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c), lo1 = 0xCD9E8D57u * c;
      a = hi1 ^ b ^ key0;
      b = lo1;
      c = hi0 ^ d ^ key1;
      d = lo0;
      key0 += 0x9E3779B9u;
      key1 += 0xBB67AE85u;
    }
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
    have = 4;
  }
  __device__ uint32_t next() {
    if (!have) refill();
    return out[--have];
  }
  // uniform integer in [0, bound)
  __device__ uint32_t below(uint32_t bound) { return __umulhi(next(), bound); }
};

struct GenParams {
  uint64_t seed;
  int total, S, n_subj, n_groups, n_cond, min_subj;
  const int *group_start;  // subject units, n_groups + 1
  int *flags;              // 1 = (re)generate this column
  int *attempts;           // draws used so far per column
  int *scratch;            // total x n_subj
  int32_t *idx;            // total x S
};

__device__ __forceinline__ int group_of(const int *gs, int n_groups, int subj) {
  int g = 0;
  while (g + 1 < n_groups && subj >= gs[g + 1]) ++g;
  return g;
}

__global__ void gen_perm_kernel(GenParams p) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= p.total || !p.flags[id]) return;
  int *perm = p.scratch + (size_t)id * p.n_subj;
  int32_t *col = p.idx + (size_t)id * p.S;
  int att = p.attempts[id];
  for (;;) {
    Philox rng(p.seed, (uint32_t)id, (uint32_t)att, 0u);
    ++att;
    for (int i = 0; i < p.n_subj; ++i) perm[i] = i;
    for (int i = p.n_subj - 1; i > 0; --i) {
      const int j = (int)rng.below((uint32_t)i + 1u);
      const int t = perm[i];
      perm[i] = perm[j];
      perm[j] = t;
    }
    bool bad = false;
    if (p.n_groups > 1) {
      for (int g = 0; g < p.n_groups && !bad; ++g) {
        const int a = p.group_start[g], b = p.group_start[g + 1];
        bool kept = true;
        for (int k = a; k < b && kept; ++k) kept = (perm[k] >= a && perm[k] < b);
        bad = kept;
      }
    }
    if (bad && att < MAX_TRIES) continue;
    // lay the column out: destination (group g, condition c, slot k)
    for (int g = 0; g < p.n_groups; ++g) {
      const int a = p.group_start[g], b = p.group_start[g + 1], ng = b - a;
      for (int k = a; k < b; ++k) {
        const int subj = perm[k];
        const int gs = group_of(p.group_start, p.n_groups, subj);
        const int sa = p.group_start[gs], sn = p.group_start[gs + 1] - sa;
        int order[MAX_COND];
        for (int c = 0; c < p.n_cond; ++c) order[c] = c;
        for (int c = p.n_cond - 1; c > 0; --c) {
          const int j = (int)rng.below((uint32_t)c + 1u);
          const int t = order[c];
          order[c] = order[j];
          order[j] = t;
        }
        for (int c = 0; c < p.n_cond; ++c)
          col[p.n_cond * a + c * ng + (k - a)] = p.n_cond * sa + order[c] * sn + (subj - sa);
      }
    }
    break;
  }
  p.attempts[id] = att;
}

__global__ void gen_boot_kernel(GenParams p) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= p.total || !p.flags[id]) return;
  int *cnt = p.scratch + (size_t)id * p.n_subj;
  int32_t *col = p.idx + (size_t)id * p.S;
  const int att = p.attempts[id];
  Philox rng(p.seed, (uint32_t)id, (uint32_t)att, 1u);
  for (int g = 0; g < p.n_groups; ++g) {
    const int a = p.group_start[g], b = p.group_start[g + 1], ng = b - a;
    for (int redo = 0; redo < 100000; ++redo) {
      for (int k = a; k < b; ++k) cnt[k] = 0;
      for (int i = 0; i < ng; ++i) cnt[a + (int)rng.below((uint32_t)ng)]++;
      int uniq = 0;
      for (int k = a; k < b; ++k) uniq += cnt[k] > 0;
      if (uniq >= p.min_subj) break;
    }
    int slot = 0;
    for (int subj = a; subj < b; ++subj)
      for (int m = 0; m < cnt[subj]; ++m, ++slot)
        for (int c = 0; c < p.n_cond; ++c)
          col[p.n_cond * a + c * ng + slot] = p.n_cond * a + c * ng + (subj - a);
  }
  p.attempts[id] = att + 1;
}

// FNV-1a style 64-bit hash of rows [a,b) of every column, one per (column, segment)
__global__ void hash_kernel(const int32_t *__restrict__ idx, int total, int S, int n_seg,
                            const int *__restrict__ seg_bounds, uint64_t *__restrict__ hashes) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total * n_seg) return;
  const int id = e / n_seg, sg = e - id * n_seg;
  const int a = seg_bounds[sg], b = seg_bounds[sg + 1];
  uint64_t hsh = 1469598103934665603ull;
  for (int s = a; s < b; ++s) {
    hsh ^= (uint64_t)(uint32_t)idx[(size_t)id * S + s];
    hsh *= 1099511628211ull;
  }
  hashes[e] = hsh;
}

// Duplicate detection through an open-addressing hash table keyed by (segment,
// hash of the segment's rows); the value of a key is the SMALLEST column index that
// carries it.  Column i repeats an earlier column iff that index is < i and the rows
// agree (the compare guards against hash collisions with the table's owner).
constexpr unsigned long long HT_EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long ht_key(uint64_t hsh, int sg) {
  unsigned long long k = hsh ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(sg + 1));
  return k == HT_EMPTY ? k - 1 : k;
}

__global__ void ht_insert_kernel(const uint64_t *__restrict__ hashes, int total, int n_seg,
                                 unsigned long long *__restrict__ keys, int *__restrict__ vals,
                                 unsigned mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total * n_seg) return;
  const int id = e / n_seg, sg = e - id * n_seg;
  const unsigned long long key = ht_key(hashes[e], sg);
  unsigned slot = (unsigned)(key * 0xD6E8FEB86659FD93ull >> 32) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&keys[slot], HT_EMPTY, key);
    if (prev == HT_EMPTY || prev == key) {
      atomicMin(&vals[slot], id);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// flags[i] = 1 when column i repeats an earlier column (in any segment) and may
// still be re-drawn; counters[0] = columns to re-draw, counters[1] = exhausted
__global__ void dup_kernel(const int32_t *__restrict__ idx, int total, int S, int n_seg,
                           const int *__restrict__ seg_bounds,
                           const uint64_t *__restrict__ hashes,
                           const unsigned long long *__restrict__ keys,
                           const int *__restrict__ vals, unsigned mask,
                           const int *__restrict__ attempts, int *__restrict__ flags,
                           int *__restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  bool dup = false;
  for (int sg = 0; sg < n_seg && !dup; ++sg) {
    const unsigned long long key = ht_key(hashes[(size_t)i * n_seg + sg], sg);
    unsigned slot = (unsigned)(key * 0xD6E8FEB86659FD93ull >> 32) & mask;
    while (keys[slot] != key) slot = (slot + 1) & mask;   // the key was inserted: it is there
    const int j = vals[slot];
    if (j < i) {
      const int a = seg_bounds[sg], b = seg_bounds[sg + 1];
      bool same = true;
      for (int s = a; s < b && same; ++s)
        same = idx[(size_t)i * S + s] == idx[(size_t)j * S + s];
      dup = same;
    }
  }
  int f = 0;
  if (dup) {
    if (attempts[i] < MAX_TRIES) {
      f = 1;
      atomicAdd(&counters[0], 1);
    } else {
      atomicAdd(&counters[1], 1);
    }
  }
  flags[i] = f;
}

// Half / half (or train / test) masks, the device counterpart of gen_splits
// (pyls/base.py:162-229).  One CTA makes the `n_split` masks of one set (the
// reference draws a fresh set per permutation, base.py:704-708, and one for
// the original data): per group a coin flip between ceil and floor of
// n_g * frac subjects (base.py:205-206), drawn without replacement by
// selection sampling; conditions follow their subject (base.py:211-215); a mask
// equal to an earlier one of the same set is re-drawn, at most 500 draws per
// mask (base.py:200-201, 217-222).  Stream: Philox keyed by (seed, set id,
// attempt, split), so a set does not depend on how the sets are sharded.
__global__ void __launch_bounds__(128)
gen_splits_kernel(uint64_t seed, long long first, int n_split, double frac, int S, int n_groups,
                  int n_cond, const int *__restrict__ group_start, int32_t *__restrict__ masks,
                  int *__restrict__ n_exhausted) {
  extern __shared__ unsigned long long sh_hash[];   // n_split hashes, then n_split redo flags
  int *redo = reinterpret_cast<int *>(sh_hash + n_split);
  const int set = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const uint32_t id = (uint32_t)(first + set);
  int32_t *out = masks + (size_t)set * n_split * S;
  for (int i = tid; i < n_split; i += nt) redo[i] = 1;
  for (int round = 0; round < MAX_TRIES; ++round) {
    __syncthreads();
    for (int i = tid; i < n_split; i += nt) {
      if (!redo[i]) continue;
      Philox rng(seed, id, (uint32_t)round, 0x10000u + (uint32_t)i);
      int32_t *col = out + (size_t)i * S;
      unsigned long long hsh = 1469598103934665603ull;
      for (int g = 0; g < n_groups; ++g) {
        const int a = group_start[g], ng = group_start[g + 1] - a;
        const double want = ng * frac;
        int need = (rng.next() & 1u) ? (int)floor(want) : (int)ceil(want);
        for (int k = 0; k < ng; ++k) {
          const int sel = (int)rng.below((uint32_t)(ng - k)) < need ? 1 : 0;
          need -= sel;
          for (int c = 0; c < n_cond; ++c) col[n_cond * a + c * ng + k] = sel;
          hsh = (hsh ^ (unsigned long long)(sel + 1)) * 1099511628211ull;
        }
      }
      sh_hash[i] = hsh;
    }
    __syncthreads();
    // a mask that repeats an earlier mask of the set is drawn again
    int dup_any = 0;
    for (int i = tid; i < n_split; i += nt) {
      bool dup = false;
      const unsigned long long hi = sh_hash[i];
      for (int j = 0; j < i && !dup; ++j) dup = sh_hash[j] == hi;
      if (dup && round + 1 >= MAX_TRIES) {
        atomicAdd(n_exhausted, 1);
        dup = false;
      }
      redo[i] = dup ? 1 : 0;
      dup_any |= dup;
    }
    // one uniform decision for the whole CTA (barrier + reduction in one)
    if (!__syncthreads_or(dup_any)) break;
  }
}

// ---- Gaussian test matrices of the randomized range finder ---------------------------
// SIMPLS resample i takes its (T, 11) test matrix from RandomState(i).normal(size=(T, 11))
// (compute.svd(Cov, n_components=1, seed=i), pyls/types/regression.py:103 ->
// sklearn randomized_svd): NumPy's legacy stream, i.e. MT19937 seeded by init_genrand(i),
// 53-bit doubles from two outputs, and the Marsaglia polar method with its cached second
// deviate.  One thread replays one stream (624 words of state in local memory): the
// uniforms are bit-exact, the deviates equal NumPy's up to the last bit of log / sqrt.
struct MT19937 {
  uint32_t key[624];
  int pos;
  __device__ void seed(uint32_t s) {
    for (int i = 0; i < 624; ++i) {
      key[i] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
    pos = 624;
  }
  __device__ void twist() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int i = 0;
    for (; i < 624 - 397; ++i) {
      const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; i < 623; ++i) {
      const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    const uint32_t y = (key[623] & UPPER) | (key[0] & LOWER);
    key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    pos = 0;
  }
  __device__ uint32_t next() {
    if (pos == 624) twist();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  __device__ double next_double() {
    const uint32_t a = next() >> 5, b = next() >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
  }
};

__global__ void __launch_bounds__(64)
gaussian_tables_kernel(long long first, int count, int per, double *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= count) return;
  MT19937 mt;
  mt.seed((uint32_t)(first + r));
  double *o = out + (size_t)r * per;
  bool has = false;
  double cached = 0.0;
  for (int e = 0; e < per; ++e) {
    if (has) {
      o[e] = cached;
      has = false;
      continue;
    }
    double x1, x2, r2;
    do {
      x1 = __dsub_rn(__dmul_rn(2.0, mt.next_double()), 1.0);
      x2 = __dsub_rn(__dmul_rn(2.0, mt.next_double()), 1.0);
      r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
    cached = __dmul_rn(f, x1);
    has = true;
    o[e] = __dmul_rn(f, x2);
  }
}

__global__ void fill_int_kernel(int *p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

int gen_indices(plsb_ctx *h, bool boot, uint64_t seed, int64_t first, int count, int32_t *d_idx,
                int *h_n_exhausted, cudaStream_t st) {
  KernelTimer kt(h, KC_INDEX, st);
  const Layout &l = h->lay;
  PLSB_CHECK(first >= 0 && count >= 0 && first + count < (1ll << 30), PLSB_ERR_ARG,
             "index generation: bad range first=%lld count=%d", (long long)first, count);
  PLSB_CHECK(l.n_cond <= MAX_COND, PLSB_ERR_ARG, "index generation: n_cond=%d > %d", l.n_cond,
             MAX_COND);
  if (h_n_exhausted) *h_n_exhausted = 0;
  if (count == 0) return PLSB_OK;
  const int total = (int)(first + count);
  const int n_seg = boot ? l.n_groups : 1;
  // segments whose equality defines a duplicate (rows; the bootstrap bounds are
  // the reference's subject-unit bounds used as row bounds)
  std::vector<int> seg(n_seg + 1);
  if (boot) {
    int acc = 0;
    for (int g = 0; g < l.n_groups; ++g) {
      seg[g] = acc;
      acc += l.groups[g];
    }
    seg[l.n_groups] = acc;
  } else {
    seg[0] = 0;
    seg[1] = l.S;
  }
  PLSB_TRY(h->idxall.ensure(sizeof(int32_t) * (size_t)total * l.S));
  // hash table: power of two >= 2 x (columns x segments)
  size_t ht_size = 1024;
  while (ht_size < 2 * (size_t)total * n_seg) ht_size <<= 1;
  const unsigned ht_mask = (unsigned)(ht_size - 1);
  const size_t n_int = (size_t)total * 2 + (size_t)total * l.n_subj + 2 + (n_seg + 1) + ht_size;
  PLSB_TRY(h->flags.ensure(sizeof(int) * n_int +
                           sizeof(uint64_t) * ((size_t)total * n_seg + 1 + ht_size)));
  uint64_t *hashes = h->flags.as<uint64_t>();
  unsigned long long *ht_keys =
      reinterpret_cast<unsigned long long *>(hashes + (size_t)total * n_seg + 1);
  int *flags = reinterpret_cast<int *>(ht_keys + ht_size);
  int *attempts = flags + total;
  int *scratch = attempts + total;
  int *counters = scratch + (size_t)total * l.n_subj;
  int *d_seg = counters + 2;
  int *ht_vals = d_seg + (n_seg + 1);
  PLSB_CUDA(cudaMemcpyAsync(d_seg, seg.data(), sizeof(int) * (n_seg + 1), cudaMemcpyHostToDevice,
                            st));
  const int tb = 128, nb = cdiv(total, tb);
  fill_int_kernel<<<nb, tb, 0, st>>>(flags, total, 1);
  PLSB_LAUNCHED(h);
  PLSB_CUDA(cudaMemsetAsync(attempts, 0, sizeof(int) * total, st));

  GenParams p;
  p.seed = seed;
  p.total = total;
  p.S = l.S;
  p.n_subj = l.n_subj;
  p.n_groups = l.n_groups;
  p.n_cond = l.n_cond;
  int min_group = l.groups[0];
  for (int g : l.groups) min_group = std::min(min_group, g);
  p.min_subj = (min_group + 1) / 2;
  p.group_start = h->d_group_start;
  p.flags = flags;
  p.attempts = attempts;
  p.scratch = scratch;
  p.idx = h->idxall.as<int32_t>();

  int host_counters[2] = {0, 0};
  for (int round = 0; round <= MAX_TRIES; ++round) {
    if (boot)
      gen_boot_kernel<<<nb, tb, 0, st>>>(p);
    else
      gen_perm_kernel<<<nb, tb, 0, st>>>(p);
    PLSB_LAUNCHED(h);
    hash_kernel<<<cdiv(total * n_seg, tb), tb, 0, st>>>(p.idx, total, l.S, n_seg, d_seg, hashes);
    PLSB_LAUNCHED(h);
    PLSB_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(int), st));
    PLSB_CUDA(cudaMemsetAsync(ht_keys, 0xFF, sizeof(unsigned long long) * ht_size, st));
    PLSB_CUDA(cudaMemsetAsync(ht_vals, 0x7F, sizeof(int) * ht_size, st));
    ht_insert_kernel<<<cdiv(total * n_seg, tb), tb, 0, st>>>(hashes, total, n_seg, ht_keys,
                                                             ht_vals, ht_mask);
    PLSB_LAUNCHED(h);
    dup_kernel<<<nb, tb, 0, st>>>(p.idx, total, l.S, n_seg, d_seg, hashes, ht_keys, ht_vals,
                                  ht_mask, attempts, flags, counters);
    PLSB_LAUNCHED(h);
    PLSB_CUDA(cudaMemcpyAsync(host_counters, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                              st));
    PLSB_CUDA(cudaStreamSynchronize(st));
    if (host_counters[0] == 0) break;
  }
  if (h_n_exhausted) *h_n_exhausted = host_counters[1];
  PLSB_CUDA(cudaMemcpyAsync(d_idx, h->idxall.as<int32_t>() + (size_t)first * l.S,
                            sizeof(int32_t) * (size_t)count * l.S, cudaMemcpyDeviceToDevice, st));
  PLSB_CUDA(cudaStreamSynchronize(st));
  return PLSB_OK;
}

int gen_gaussian_tables(plsb_ctx *h, int64_t first, int count, int per, double *d_out,
                        cudaStream_t st) {
  KernelTimer kt(h, KC_INDEX, st);
  PLSB_CHECK(first >= 0 && count >= 0 && first + count <= 0xffffffffll && per >= 1, PLSB_ERR_ARG,
             "Gaussian tables: seeds must lie in [0, 2^32)");
  if (count == 0) return PLSB_OK;
  gaussian_tables_kernel<<<cdiv(count, 64), 64, 0, st>>>(first, count, per, d_out);
  PLSB_LAUNCHED(h);
  return PLSB_OK;
}

int gen_split_masks(plsb_ctx *h, uint64_t seed, int64_t first, int count, int n_split, double frac,
                    int32_t *d_masks, int *h_n_exhausted, cudaStream_t st) {
  KernelTimer kt(h, KC_INDEX, st);
  const Layout &l = h->lay;
  PLSB_CHECK(first >= 0 && count >= 0 && first + count < (1ll << 30) && n_split >= 1 &&
                 n_split < 0x10000 && frac > 0.0 && frac < 1.0,
             PLSB_ERR_ARG, "split generation: bad argument");
  if (h_n_exhausted) *h_n_exhausted = 0;
  if (count == 0) return PLSB_OK;
  PLSB_TRY(h->flags.ensure(sizeof(int)));
  int *counter = h->flags.as<int>();
  PLSB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  gen_splits_kernel<<<count, 128, (sizeof(unsigned long long) + sizeof(int)) * n_split, st>>>(
      seed, first, n_split, frac, l.S, l.n_groups, l.n_cond, h->d_group_start, d_masks, counter);
  PLSB_LAUNCHED(h);
  if (h_n_exhausted) {
    PLSB_CUDA(cudaMemcpyAsync(h_n_exhausted, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    PLSB_CUDA(cudaStreamSynchronize(st));
  }
  return PLSB_OK;
}

}  // namespace plsb
