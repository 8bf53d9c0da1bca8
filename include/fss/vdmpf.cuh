// SPDX-License-Identifier: Apache-2.0
// fss/vdmpf.cuh -- 2-party verifiable distributed MULTI-point function (reference vdmpf.cuh:80-279; de Castro &
// Polychroniadou, ePrint 2024/677 section 4): same class template, template parameter list, `BucketKey` / `Key` layouts and
// member signatures.  The t points are cuckoo-hashed (fss/cuckoo_hash.cuh) into m buckets; every bucket is one inner VDPF
// over a `bucket_bits` domain, and an input is looked up in its kappa candidate buckets.
//
// What runs where.  The reference walks the buckets one inner `Gen` / `Eval` / hash at a time on the CPU.  Here the inner
// VDPFs are a BATCH for the B200:
//   * Gen: the cuckoo table is built on the host (a few PRP calls per point), then the m inner keys come from ONE
//     fssb200_vdpf_gen_host call (m keys, one kernel);
//   * BatchEval: the kappa places of every input are computed on the host, the (bucket, position) pairs are grouped by
//     bucket in the reference's order, and ALL inner evaluations of the call -- up to kappa per input -- go through ONE
//     fssb200_vdpf_eval_host call; the per-bucket proof chains run as fssb200_vdpf_prove batches (the buckets that hold the
//     same number of inputs form one batch) and the chain over the m buckets is one more such call.
// Outputs and proofs are bit-identical to the reference's for the same key and inputs (tests/test_vdmpf.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstring>
#include <span>
#include <type_traits>
#include <vector>
#include <cuda/std/array>
#include <cuda/std/span>
#include <fss/cuckoo_hash.cuh>
#include <fss/group.cuh>
#include <fss/hash.cuh>
#include <fss/prg.cuh>
#include <fss/prp.cuh>
#include <fss/util.cuh>
#include <fss/vdpf.cuh>

namespace fss {

template <int in_bits, int max_points, int bucket_bits, typename Group, typename Prg, typename XorHash, typename Hash,
    typename Prp, typename In = uint, int kappa = 3, int ch_lambda = 80>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 && Groupable<Group> &&
           XorHashable<XorHash> && Hashable<Hash> && Permutable<Prp>)
class Vdmpf {
public:
  static_assert(max_points >= 30, "max_points must be >= 30 (Remark 1 of the paper)");
  static constexpr int m = cuckoo_hash::ChBucket(max_points, ch_lambda);  // buckets the key types are sized for
  static constexpr __uint128_t n = __uint128_t(1) << in_bits;
  static constexpr int b_size = static_cast<int>((n * kappa + m - 1) / m);  // values of [0, kappa n) per bucket
  static_assert(b_size <= (1 << bucket_bits));

  using InnerVdpf = Vdpf<bucket_bits, Group, Prg, XorHash, Hash, uint>;
  using InnerCw = typename InnerVdpf::Cw;

  Prg prg;
  XorHash xor_hash;
  Hash hash;
  Prp prp;

  struct BucketKey {  // vdmpf.cuh:104-109
    InnerCw cws[bucket_bits];
    cuda::std::array<int4, 4> cs;
    int4 ocw;
    int4 s0;
  };
  struct Key {  // vdmpf.cuh:116-121: what Gen chose at run time travels with the key
    int4 sigma;
    int m_rt;
    int b_size_rt;
    BucketKey bks[m];
  };

  // vdmpf.cuh:136: 0, or 1 when the cuckoo hashing or an inner VDPF Gen failed (draw new sigma / seeds and call again).
  int Gen(Key &k0, Key &k1, int4 sigma, cuda::std::span<const cuda::std::array<int4, 2>, m> s0s, std::span<const In> as,
      std::span<const int4> b_bufs, int t, int ch_retry = 1000) {
    assert(t >= 1 && t <= max_points);  // (the paper's failure bound is stated for t >= 30; the reference's own sample hashes 8 points)
    const int m_rt = cuckoo_hash::ChBucket(t, ch_lambda);
    const int b_rt = static_cast<int>((n * kappa + m_rt - 1) / m_rt);
    assert(m_rt <= m && b_rt <= (1 << bucket_bits));
    k0.sigma = k1.sigma = sigma;
    k0.m_rt = k1.m_rt = m_rt;
    k0.b_size_rt = k1.b_size_rt = b_rt;

    std::vector<std::pair<int, int>> table(size_t(m_rt), {-1, -1});
    cuckoo_hash::Compact<Prp, In, kappa> compact{prp};
    if (compact.Run(as.first(size_t(t)), m_rt, sigma, n, b_rt, ch_retry, std::span<std::pair<int, int>>(table)) != 0) return 1;

    // The inner point functions: bucket i holds (position of its point inside the bucket, beta), or the zero function.
    cuckoo_hash::PrpHash<Prp, In, kappa> where{prp};
    std::vector<uint> alphas(static_cast<size_t>(m), 0u);
    std::vector<int4> betas(static_cast<size_t>(m), int4{0, 0, 0, 0});
    for (int i = 0; i < m_rt; ++i) {
      const auto [j, k] = table[size_t(i)];
      if (j < 0) continue;
      alphas[size_t(i)] = static_cast<uint>(where.Locate(sigma, as[size_t(j)], k, n, b_rt).second);
      assert(alphas[size_t(i)] < (1u << bucket_bits));
      betas[size_t(i)] = b_bufs[size_t(j)];
    }
    // ... generated as one batch of m keys
    const size_t nb = m;
    std::vector<InnerCw> cws(nb * bucket_bits);
    std::vector<int4> cs(nb * 4), ocws(nb);
    std::vector<int32_t> status(nb, 0);
    InnerVdpf inner{prg, xor_hash, hash};
    b200::Check(fssb200_vdpf_gen_host(inner.Context(), s0s.data(), alphas.data(), betas.data(), cws.data(), cs.data(),
                    ocws.data(), status.data(), size_t(m)),
        "Vdmpf::Gen");
    for (int i = 0; i < m; ++i) {
      if (status[size_t(i)] != 0) return 1;
      for (Key *k : {&k0, &k1}) {
        BucketKey &bk = k->bks[i];
        std::memcpy(bk.cws, &cws[size_t(i) * bucket_bits], sizeof(bk.cws));
        std::memcpy(bk.cs.data(), &cs[size_t(i) * 4], 64);
        bk.ocw = ocws[size_t(i)];
      }
      k0.bks[i].s0 = s0s[size_t(i)][0];
      k1.bks[i].s0 = s0s[size_t(i)][1];
    }
    return 0;
  }

  // vdmpf.cuh:196: output shares of xs (ys zero where no bucket holds x) and the proof over everything that was evaluated.
  void BatchEval(bool b, const Key &key, std::span<const In> xs, std::span<int4> ys, cuda::std::array<int4, 4> &pi) {
    const size_t eta = xs.size();
    assert(ys.size() >= eta);
    const int b_rt = key.b_size_rt;

    // 1. every (bucket, position, input) the call has to evaluate, bucket by bucket, inside a bucket in input order
    //    (the order the proof chain of a bucket absorbs them in, vdmpf.cuh:207-228)
    struct Visit {
      int bucket;
      uint pos;
      size_t input;
    };
    cuckoo_hash::PrpHash<Prp, In, kappa> where{prp};
    const std::vector<std::pair<int, int>> places = where.Locations(key.sigma, xs, n, b_rt);
    std::vector<Visit> visits;
    visits.reserve(places.size());
    std::vector<size_t> first(size_t(m) + 1, 0);  // visits of bucket i: [first[i], first[i + 1])
    for (size_t w = 0; w < eta; ++w)
      for (int k = 0; k < kappa; ++k) {
        const auto [bucket, pos] = places[w * size_t(kappa) + size_t(k)];
        if (bucket >= m) continue;
        bool twice = false;  // (two hash functions cannot agree on a place -- the PRP is injective -- but the rule is kept)
        for (int k2 = 0; k2 < k; ++k2) twice |= places[w * size_t(kappa) + size_t(k2)] == places[w * size_t(kappa) + size_t(k)];
        if (twice) continue;
        visits.push_back({bucket, static_cast<uint>(pos), w});
        ++first[size_t(bucket) + 1];
      }
    for (int i = 0; i < m; ++i) first[size_t(i) + 1] += first[size_t(i)];
    {  // stable counting sort by bucket
      std::vector<Visit> sorted(visits.size());
      std::vector<size_t> at(first.begin(), first.end() - 1);
      for (const Visit &v : visits) sorted[at[size_t(v.bucket)]++] = v;
      visits.swap(sorted);
    }
    const size_t nv = visits.size();

    // 2. one batch of inner evaluations: visit e evaluates bucket key visits[e].bucket at visits[e].pos
    std::vector<int4> seeds(nv), cs(nv * 4), ocws(nv), shares(nv), pit(nv * 4);
    std::vector<InnerCw> cws(nv * bucket_bits);
    std::vector<uint> pos(nv);
    for (size_t e = 0; e < nv; ++e) {
      const BucketKey &bk = key.bks[visits[e].bucket];
      seeds[e] = bk.s0;
      std::memcpy(&cws[e * bucket_bits], bk.cws, sizeof(bk.cws));
      std::memcpy(&cs[e * 4], bk.cs.data(), 64);
      ocws[e] = bk.ocw;
      pos[e] = visits[e].pos;
    }
    InnerVdpf inner{prg, xor_hash, hash};
    fssb200_ctx *ctx = inner.Context();
    if (nv)
      b200::Check(fssb200_vdpf_eval_host(ctx, b, seeds.data(), cws.data(), cs.data(), ocws.data(), pos.data(), shares.data(),
                      pit.data(), nv),
          "Vdmpf::BatchEval");

    // 3. an input's share is the sum of its buckets' shares
    for (size_t w = 0; w < eta; ++w) ys[w] = int4{0, 0, 0, 0};
    for (size_t e = 0; e < nv; ++e) ys[visits[e].input] = (Group::From(ys[visits[e].input]) + Group::From(shares[e])).Into();

    // 4. proofs.  Bucket i: the chain of Vdpf::Prove over its visits, started from its cs.  Buckets with the same number
    //    of visits c > 0 are one fssb200_vdpf_prove batch [buckets][c][4]; then one chain over the m bucket proofs from zero.
    std::vector<int4> bucket_pi(size_t(m) * 4);
    for (int i = 0; i < m; ++i) std::memcpy(&bucket_pi[size_t(i) * 4], key.bks[i].cs.data(), 64);
    size_t most = 0;
    for (int i = 0; i < m; ++i) most = std::max(most, first[size_t(i) + 1] - first[size_t(i)]);
    if (nv) {
      b200::DeviceArray<int4> d_pit(nv * 4), d_cs(size_t(m) * 4), d_pi(size_t(m) * 4);
      std::vector<int4> g_pit, g_cs, g_pi;
      std::vector<int> members;
      for (size_t c = 1; c <= most; ++c) {
        members.clear();
        g_pit.clear();
        g_cs.clear();
        for (int i = 0; i < m; ++i)
          if (first[size_t(i) + 1] - first[size_t(i)] == c) {
            members.push_back(i);
            g_pit.insert(g_pit.end(), pit.begin() + ptrdiff_t(first[size_t(i)] * 4), pit.begin() + ptrdiff_t(first[size_t(i) + 1] * 4));
            g_cs.insert(g_cs.end(), bucket_pi.begin() + ptrdiff_t(i) * 4, bucket_pi.begin() + ptrdiff_t(i) * 4 + 4);
          }
        if (members.empty()) continue;
        d_pit.Upload(0, g_pit.data(), g_pit.size());
        d_cs.Upload(0, g_cs.data(), g_cs.size());
        b200::Check(fssb200_vdpf_prove(ctx, d_pit.ptr, d_cs.ptr, c, d_pi.ptr, members.size(), nullptr), "Vdmpf::BatchEval (bucket proofs)");
        g_pi.resize(members.size() * 4);
        d_pi.Download(0, g_pi.data(), g_pi.size());
        for (size_t g = 0; g < members.size(); ++g) std::memcpy(&bucket_pi[size_t(members[g]) * 4], &g_pi[g * 4], 64);
      }
    }
    {
      b200::DeviceArray<int4> d(size_t(m) * 4 + 8);
      const int4 zero[4] = {};
      d.Upload(0, bucket_pi.data(), size_t(m) * 4);
      d.Upload(size_t(m) * 4, zero, 4);
      b200::Check(fssb200_vdpf_prove(ctx, d.ptr, d.ptr + size_t(m) * 4, size_t(m), d.ptr + size_t(m) * 4 + 4, 1, nullptr),
          "Vdmpf::BatchEval (proof over the buckets)");
      d.Download(size_t(m) * 4 + 4, pi.data(), 4);
    }
  }

  // vdmpf.cuh:274
  static bool Verify(cuda::std::span<const int4, 4> pi0, cuda::std::span<const int4, 4> pi1) {
    return InnerVdpf::Verify(pi0, pi1);
  }
};

}  // namespace fss
