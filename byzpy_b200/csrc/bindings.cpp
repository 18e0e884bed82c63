// pybind11 bindings for the byzpy_b200 kernel library.
//
// Deliberately torch-free: tensors cross the boundary as raw device addresses
// (tensor.data_ptr()) and the CUDA stream as an integer handle
// (torch.cuda.current_stream().cuda_stream), so this file compiles in seconds
// and the .so has no libtorch ABI dependency.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "api.h"
#include <algorithm>
#include <stdexcept>

#include "bn.h"
#include "pool.h"
#include "layout.h"
#include "gram_umma.h"
#include "nspace.h"
#include "runtime.h"
#include "host_select.h"

namespace py = pybind11;

namespace {

void check(int err, const char* what) {
  if (err != 0) {
    throw std::runtime_error(std::string(what) + ": CUDA error " + std::to_string(err) + " (" +
                             cudaGetErrorString((cudaError_t)err) + ")");
  }
}

template <typename T>
T* as_ptr(uint64_t v) {
  return reinterpret_cast<T*>(static_cast<uintptr_t>(v));
}

void fill_rows(RowTable& rt, ScaleTable& st, const std::vector<uint64_t>& rows,
               const std::vector<float>& scales) {
  if (rows.empty() || rows.size() > BZ_MAXN) throw std::invalid_argument("need 1..128 rows");
  if (!scales.empty() && scales.size() != rows.size())
    throw std::invalid_argument("scales must match rows");
  for (size_t i = 0; i < BZ_MAXN; ++i) {
    rt.p[i] = i < rows.size() ? as_ptr<const float>(rows[i]) : nullptr;
    st.s[i] = (i < rows.size() && !scales.empty()) ? scales[i] : 1.0f;
  }
}

void fill_upd(UpdTable& u, const std::vector<uint64_t>& params, const std::vector<uint64_t>& moms,
              float lr, float mu, float wd) {
  if (params.size() > BZ_MAXR) throw std::invalid_argument("too many replicas (max 16)");
  if (!moms.empty() && moms.size() != params.size())
    throw std::invalid_argument("moms must match params");
  std::memset(&u, 0, sizeof(u));
  u.count = (int)params.size();
  for (size_t r = 0; r < params.size(); ++r) {
    u.param[r] = as_ptr<float>(params[r]);
    u.mom[r] = moms.empty() ? nullptr : as_ptr<float>(moms[r]);
  }
  u.lr = lr;
  u.mu = mu;
  u.wd = wd;
}

cudaStream_t as_stream(uint64_t s) { return reinterpret_cast<cudaStream_t>(static_cast<uintptr_t>(s)); }

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.doc() = "byzpy_b200 sm_100a kernel library";
  m.attr("MAXN") = BZ_MAXN;
  m.attr("MAXR") = BZ_MAXR;
  m.attr("ARCH") = "sm_100a";

  m.def(
      "cw_select",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, int mode, int f,
         int n_virtual, int n_honest, float va, float vb, long long off, long long len,
         uint64_t out, const std::vector<uint64_t>& upd_params,
         const std::vector<uint64_t>& upd_moms, float lr, float mu, float wd, int sm_count,
         uint64_t stream, int impl) {
        BzCwArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.virt.count = n_virtual;
        a.virt.n_honest = n_honest;
        a.virt.a = va;
        a.virt.b = vb;
        a.f = f;
        a.mode = mode;
        a.off = off;
        a.len = len;
        a.out = as_ptr<float>(out);
        fill_upd(a.upd, upd_params, upd_moms, lr, mu, wd);
        a.impl = impl;
        check(bz_cw_select(&a, sm_count, as_stream(stream)), "cw_select");
      },
      py::arg("rows"), py::arg("scales"), py::arg("mode"), py::arg("f"), py::arg("n_virtual"),
      py::arg("n_honest"), py::arg("va"), py::arg("vb"), py::arg("off"), py::arg("len"),
      py::arg("out"), py::arg("upd_params"), py::arg("upd_moms"), py::arg("lr"), py::arg("mu"),
      py::arg("wd"), py::arg("sm_count"), py::arg("stream"), py::arg("impl") = 0);

  m.def(
      "wsum",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, uint64_t W, int mrows,
         long long off, long long len, const std::vector<uint64_t>& outs,
         const std::vector<uint64_t>& upd_params, const std::vector<uint64_t>& upd_moms, float lr,
         float mu, float wd, int sm_count, uint64_t stream) {
        BzWsumArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.m = mrows;
        a.W = as_ptr<const float>(W);
        a.off = off;
        a.len = len;
        if ((int)outs.size() != mrows || mrows > 8) throw std::invalid_argument("outs must have m<=8 entries");
        for (int r = 0; r < mrows; ++r) a.out[r] = as_ptr<float>(outs[r]);
        fill_upd(a.upd, upd_params, upd_moms, lr, mu, wd);
        check(bz_wsum(&a, sm_count, as_stream(stream)), "wsum");
      },
      py::arg("rows"), py::arg("scales"), py::arg("W"), py::arg("m"), py::arg("off"),
      py::arg("len"), py::arg("outs"), py::arg("upd_params"), py::arg("upd_moms"), py::arg("lr"),
      py::arg("mu"), py::arg("wd"), py::arg("sm_count"), py::arg("stream"));

  m.attr("WSUM_MULTI_TILE") = bz_wsum_multi_tile();
  m.def(
      "wsum_multi",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, uint64_t W, int mrows, long long off,
         long long len, const std::vector<uint64_t>& outs, int sm_count, uint64_t stream) {
        BzWsumMultiArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.m = mrows;
        a.W = as_ptr<const float>(W);
        a.off = off;
        a.len = len;
        if ((int)outs.size() != mrows || mrows > BZ_MAXN) throw std::invalid_argument("outs must have m <= 128 entries");
        for (int r = 0; r < mrows; ++r) a.out.p[r] = as_ptr<float>(outs[r]);
        check(bz_wsum_multi(&a, sm_count, as_stream(stream)), "wsum_multi");
      },
      py::arg("rows"), py::arg("scales"), py::arg("W"), py::arg("m"), py::arg("off"), py::arg("len"),
      py::arg("outs"), py::arg("sm_count"), py::arg("stream"));

  m.def(
      "host_colstat",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, double a, double b, long long d,
         uint64_t out, int threads) {
        if (rows.empty()) throw std::invalid_argument("need at least one row");
        if (!scales.empty() && scales.size() != rows.size())
          throw std::invalid_argument("scales must match rows");
        std::vector<const float*> ptrs(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) ptrs[i] = as_ptr<const float>(rows[i]);
        int rc;
        {
          py::gil_scoped_release nogil;
          rc = bz_host_colstat(ptrs.data(), scales.empty() ? nullptr : scales.data(), (int)rows.size(),
                               (int64_t)d, a, b, as_ptr<float>(out), threads);
        }
        if (rc != 0) throw std::invalid_argument("host_colstat: bad arguments");
      },
      py::arg("rows"), py::arg("scales"), py::arg("a"), py::arg("b"), py::arg("d"), py::arg("out"),
      py::arg("threads"));
  m.def("host_network_size", &bz_host_network_size);
  m.def(
      "host_cw_select",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, int mode, int f, long long d,
         uint64_t out, int threads) {
        if (rows.empty()) throw std::invalid_argument("need at least one row");
        if (!scales.empty() && scales.size() != rows.size())
          throw std::invalid_argument("scales must match rows");
        std::vector<const float*> ptrs(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) ptrs[i] = as_ptr<const float>(rows[i]);
        int rc;
        {
          py::gil_scoped_release nogil;   // pure host compute on caller-owned buffers
          rc = bz_host_cw_select(ptrs.data(), scales.empty() ? nullptr : scales.data(), (int)rows.size(),
                                 (int64_t)d, mode, f, as_ptr<float>(out), threads);
        }
        if (rc != 0) throw std::invalid_argument("host_cw_select: bad arguments (code " + std::to_string(rc) + ")");
      },
      py::arg("rows"), py::arg("scales"), py::arg("mode"), py::arg("f"), py::arg("d"), py::arg("out"),
      py::arg("threads"));
  m.def("gram_partials_needed", &bz_gram_partials_needed);
  m.def(
      "gram",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, long long off,
         long long len, uint64_t partials, int num_partials, uint64_t G, uint64_t G64, int sm_count,
         uint64_t stream, uint64_t aux_median) {
        BzGramArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.off = off;
        a.len = len;
        a.partials = as_ptr<float>(partials);
        a.num_partials = num_partials;
        a.G = as_ptr<float>(G);
        a.G64 = as_ptr<double>(G64);
        a.aux_median = as_ptr<float>(aux_median);
        check(bz_gram(&a, sm_count, as_stream(stream)), "gram");
      },
      py::arg("rows"), py::arg("scales"), py::arg("off"), py::arg("len"), py::arg("partials"),
      py::arg("num_partials"), py::arg("G"), py::arg("G64"), py::arg("sm_count"),
      py::arg("stream"), py::arg("aux_median") = 0);

  m.def(
      "colstat",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, float a_, float b_,
         long long off, long long len, uint64_t out, int sm_count, uint64_t stream) {
        BzColStatArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.a = a_;
        a.b = b_;
        a.off = off;
        a.len = len;
        a.out = as_ptr<float>(out);
        check(bz_colstat(&a, sm_count, as_stream(stream)), "colstat");
      },
      py::arg("rows"), py::arg("scales"), py::arg("a"), py::arg("b"), py::arg("off"),
      py::arg("len"), py::arg("out"), py::arg("sm_count"), py::arg("stream"));

  m.def("u8_affine", [](uint64_t in, uint64_t out, long long n, int C, const std::vector<float>& mean,
                        const std::vector<float>& scale, int sm_count, uint64_t stream) {
    if ((int)mean.size() < C || (int)scale.size() < C) throw std::invalid_argument("u8_affine: mean/scale");
    check(bz_u8_affine(as_ptr<const void>(in), as_ptr<void>(out), n, C, mean.data(), scale.data(), sm_count,
                       as_stream(stream)),
          "u8_affine");
  });
  m.def("scale_copy", [](uint64_t src, uint64_t dst, float scale, long long len, int sm_count,
                         uint64_t stream) {
    check(bz_scale_copy(as_ptr<const float>(src), as_ptr<float>(dst), scale, len, sm_count,
                        as_stream(stream)),
          "scale_copy");
  });
  m.def("fill", [](uint64_t dst, float value, long long len, int sm_count, uint64_t stream) {
    check(bz_fill(as_ptr<float>(dst), value, len, sm_count, as_stream(stream)), "fill");
  });
  m.def("gaussian", [](uint64_t dst, float mu, float sigma, unsigned long long seed,
                       unsigned long long offset, long long len, int sm_count, uint64_t stream) {
    check(bz_gaussian(as_ptr<float>(dst), mu, sigma, seed, offset, len, sm_count,
                      as_stream(stream)),
          "gaussian");
  });
  m.def("sgd", [](uint64_t grad, const std::vector<uint64_t>& params,
                  const std::vector<uint64_t>& moms, float lr, float mu, float wd, long long len,
                  int sm_count, uint64_t stream) {
    UpdTable u;
    fill_upd(u, params, moms, lr, mu, wd);
    check(bz_sgd(as_ptr<const float>(grad), &u, len, sm_count, as_stream(stream)), "sgd");
  });

  m.def("bn_partial_blocks", &bz_bn_partial_blocks);
  m.def(
      "bn_forward",
      [](uint64_t x, uint64_t y, long long R, int C, uint64_t gamma, uint64_t beta, uint64_t rmean,
         uint64_t rvar, uint64_t mean, uint64_t invstd, uint64_t scale, uint64_t shift, uint64_t partial,
         float eps, float momentum, int relu, int training, uint64_t res, uint64_t nbt, int sm_count,
         uint64_t stream) {
        BzBnArgs a;
        std::memset(&a, 0, sizeof(a));
        a.x = as_ptr<const void>(x);
        a.y = as_ptr<void>(y);
        a.R = R;
        a.C = C;
        a.gamma = as_ptr<const float>(gamma);
        a.beta = as_ptr<const float>(beta);
        a.running_mean = as_ptr<float>(rmean);
        a.running_var = as_ptr<float>(rvar);
        a.mean = as_ptr<float>(mean);
        a.invstd = as_ptr<float>(invstd);
        a.scale = as_ptr<float>(scale);
        a.shift = as_ptr<float>(shift);
        a.partial = as_ptr<float>(partial);
        a.eps = eps;
        a.momentum = momentum;
        a.relu = relu;
        a.training = training;
        a.res = as_ptr<const void>(res);
        a.num_batches_tracked = as_ptr<long long>(nbt);
        int launches = 0;
        a.launches = &launches;
        check(bz_bn_forward(&a, sm_count, as_stream(stream)), "bn_forward");
        return launches;
      });
  m.def(
      "bn_backward",
      [](uint64_t x, uint64_t dy, uint64_t dx, long long R, int C, uint64_t gamma, uint64_t mean,
         uint64_t invstd, uint64_t scale, uint64_t shift, uint64_t partial, uint64_t dgamma,
         uint64_t dbeta, uint64_t coef, int relu, uint64_t ymask, uint64_t dres, int sm_count,
         uint64_t stream) {
        BzBnArgs a;
        std::memset(&a, 0, sizeof(a));
        a.x = as_ptr<const void>(x);
        a.dy = as_ptr<const void>(dy);
        a.dx = as_ptr<void>(dx);
        a.R = R;
        a.C = C;
        a.gamma = as_ptr<const float>(gamma);
        a.mean = as_ptr<float>(mean);
        a.invstd = as_ptr<float>(invstd);
        a.scale = as_ptr<float>(scale);
        a.shift = as_ptr<float>(shift);
        a.partial = as_ptr<float>(partial);
        a.dgamma = as_ptr<float>(dgamma);
        a.dbeta = as_ptr<float>(dbeta);
        a.coef = as_ptr<float>(coef);
        a.relu = relu;
        a.ymask = as_ptr<const void>(ymask);
        a.dres = as_ptr<void>(dres);
        int launches = 0;
        a.launches = &launches;
        check(bz_bn_backward(&a, sm_count, as_stream(stream)), "bn_backward");
        return launches;
      });
  m.def(
      "maxpool_forward",
      [](uint64_t x, uint64_t y, uint64_t idx, int N, int H, int W, int C, int sm_count, uint64_t stream) {
        check(bz_maxpool3x3s2_forward(as_ptr<const void>(x), as_ptr<void>(y), as_ptr<void>(idx), N, H, W, C,
                                      sm_count, as_stream(stream)),
              "maxpool_forward");
      });
  m.def(
      "maxpool_backward",
      [](uint64_t dy, uint64_t idx, uint64_t dx, int N, int H, int W, int C, int sm_count, uint64_t stream) {
        check(bz_maxpool3x3s2_backward(as_ptr<const void>(dy), as_ptr<const void>(idx), as_ptr<void>(dx), N, H,
                                       W, C, sm_count, as_stream(stream)),
              "maxpool_backward");
      });
  m.def(
      "krsc_cast",
      [](const std::vector<uint64_t>& src, const std::vector<uint64_t>& dst, const std::vector<int>& K,
         const std::vector<int>& C, const std::vector<int>& RS, int to_grad, uint64_t stream) {
        const size_t n = src.size();
        if (dst.size() != n || K.size() != n || C.size() != n || RS.size() != n)
          throw std::invalid_argument("krsc_cast: list lengths differ");
        int launches = 0;
        for (size_t base = 0; base < n; base += BZ_CAST_MAX) {
          BzCastTable t;
          std::memset(&t, 0, sizeof(t));
          int total = 0;
          const size_t end = std::min(n, base + (size_t)BZ_CAST_MAX);
          for (size_t i = base; i < end; ++i) {
            BzCastEntry& e = t.e[i - base];
            e.src = as_ptr<const void>(src[i]);
            e.dst = as_ptr<void>(dst[i]);
            e.K = K[i];
            e.C = C[i];
            e.RS = RS[i];
            e.cta_start = total;
            total += bz_krsc_cast_ctas(K[i], C[i], RS[i]);
          }
          t.count = (int)(end - base);
          check(bz_krsc_cast(&t, to_grad, as_stream(stream)), "krsc_cast");
          ++launches;
        }
        return launches;
      });
  m.def("s2d_pack", [](uint64_t x, int x_is_u8, uint64_t out, int N, int H, int W, const std::vector<float>& mean,
                       const std::vector<float>& scale, uint64_t stream) {
    if (mean.size() < 3 || scale.size() < 3) throw std::invalid_argument("s2d_pack: mean/scale need 3 entries");
    check(bz_s2d_pack(as_ptr<const void>(x), x_is_u8, as_ptr<void>(out), N, H, W, mean.data(), scale.data(),
                      as_stream(stream)),
          "s2d_pack");
  });
  m.def("stem_weight_pack", [](uint64_t w, uint64_t wp, int K, uint64_t stream) {
    check(bz_stem_weight_pack(as_ptr<const float>(w), as_ptr<void>(wp), K, as_stream(stream)), "stem_weight_pack");
  });
  m.def("stem_grad_unpack", [](uint64_t gp, uint64_t g, int K, uint64_t stream) {
    check(bz_stem_grad_unpack(as_ptr<const void>(gp), as_ptr<float>(g), K, as_stream(stream)), "stem_grad_unpack");
  });
  // segments: [(base pointer of the segment's first row, rows, row stride in bytes)], cols = row length in
  // elements -> opaque bytes of a BzGramTmaMaps (tensor maps are plain data: cache and reuse them)
  m.def("gram_tma_maps", [](const std::vector<std::tuple<uint64_t, int, unsigned long long>>& segments,
                            unsigned long long cols, int n) {
    if (segments.empty() || segments.size() > BZ_GRAM_MAXSEG) throw std::invalid_argument("1..12 segments");
    BzGramTmaMaps tm;
    std::memset(&tm, 0, sizeof(tm));
    const int tile_cols = bz_gram_umma_tile_cols(n);
    int row0 = 0;
    for (size_t k = 0; k < segments.size(); ++k) {
      const auto& [base, rows, stride] = segments[k];
      if (rows < 1 || rows > 256 || (base % 16) != 0 || (stride % 16) != 0)
        throw std::invalid_argument("segment rows must be 1..256, base and stride multiples of 16 bytes");
      const unsigned long long st = rows > 1 ? stride : ((cols * 4 + 15) / 16 * 16);
      int e = bz_encode_map_2d(&tm.maps[k], as_ptr<const void>(base), (unsigned long long)rows, cols, st,
                               (unsigned)rows, (unsigned)tile_cols);
      if (e != 0) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string(e) + ")");
      tm.seg_row0[k] = row0;
      tm.seg_rows[k] = rows;
      row0 += rows;
    }
    if (row0 != n) throw std::invalid_argument("segments do not add up to n rows");
    tm.nseg = (int)segments.size();
    return py::bytes(reinterpret_cast<const char*>(&tm), sizeof(tm));
  });
  m.def(
      "gram_umma_tma",
      [](const py::bytes& maps, const std::vector<uint64_t>& rows, const std::vector<float>& scales, long long off,
         long long len, uint64_t partials, int num_partials, uint64_t tail64, uint64_t G, uint64_t G64,
         int sm_count, uint64_t stream) {
        const std::string blob = maps;
        if (blob.size() != sizeof(BzGramTmaMaps)) throw std::invalid_argument("bad tensor-map blob");
        alignas(64) BzGramTmaMaps tm;
        std::memcpy(&tm, blob.data(), sizeof(tm));
        BzGramUmmaArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.off = off;
        a.len = len;
        a.partials = as_ptr<float>(partials);
        a.num_partials = num_partials;
        a.tail64 = as_ptr<const double>(tail64);
        a.G = as_ptr<float>(G);
        a.G64 = as_ptr<double>(G64);
        check(bz_gram_umma_tma(&a, &tm, sm_count, as_stream(stream)), "gram_umma_tma");
      },
      py::arg("maps"), py::arg("rows"), py::arg("scales"), py::arg("off"), py::arg("len"), py::arg("partials"),
      py::arg("num_partials"), py::arg("tail64"), py::arg("G"), py::arg("G64"), py::arg("sm_count"),
      py::arg("stream"));
  m.def("gram_umma_grid", &bz_gram_umma_grid);
  m.def("gram_umma_tile_cols", &bz_gram_umma_tile_cols);
  m.def("gram_umma_partials", &bz_gram_umma_partials);
  m.def(
      "gram_umma",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, long long off,
         long long len, uint64_t partials, int num_partials, uint64_t tail64, uint64_t G,
         uint64_t G64, int sm_count, uint64_t stream) {
        BzGramUmmaArgs a;
        std::memset(&a, 0, sizeof(a));
        fill_rows(a.rows, a.scales, rows, scales);
        a.n = (int)rows.size();
        a.off = off;
        a.len = len;
        a.partials = as_ptr<float>(partials);
        a.num_partials = num_partials;
        a.tail64 = as_ptr<const double>(tail64);
        a.G = as_ptr<float>(G);
        a.G64 = as_ptr<double>(G64);
        check(bz_gram_umma(&a, sm_count, as_stream(stream)), "gram_umma");
      },
      py::arg("rows"), py::arg("scales"), py::arg("off"), py::arg("len"), py::arg("partials"),
      py::arg("num_partials"), py::arg("tail64"), py::arg("G"), py::arg("G64"),
      py::arg("sm_count"), py::arg("stream"));
  m.def("binomial", &bz_binomial);
  m.def("nspace_subset_blocks", &bz_nspace_subset_blocks);
  m.def("nspace_subset", [](uint64_t G, int ldg, int n, int m_, int nt, int mode, uint64_t sscore, uint64_t srank,
                            uint64_t w, int sm_count, uint64_t stream) {
    check(bz_nspace_subset(as_ptr<const double>(G), ldg, n, m_, nt, mode, as_ptr<double>(sscore),
                           as_ptr<unsigned long long>(srank), as_ptr<float>(w), sm_count, as_stream(stream)),
          "nspace_subset");
  });
  m.def("nspace_krum", [](uint64_t G, int n, int f, int q, uint64_t w, uint64_t stream) {
    check(bz_nspace_krum(as_ptr<const double>(G), n, f, q, as_ptr<float>(w), as_stream(stream)),
          "nspace_krum");
  });
  m.def("nspace_weiszfeld", [](uint64_t G, int nt, int n_real, uint64_t a0, double tol, int max_iter,
                               double eps, uint64_t out, uint64_t iters, uint64_t stream) {
    check(bz_nspace_weiszfeld(as_ptr<const double>(G), nt, n_real, as_ptr<const double>(a0), tol,
                              max_iter, eps, as_ptr<float>(out), as_ptr<int>(iters),
                              as_stream(stream)),
          "nspace_weiszfeld");
  });
  m.def("nspace_cclip", [](uint64_t G, int nt, int n_real, uint64_t a0, double c_tau, int M,
                           double eps, uint64_t out, uint64_t stream) {
    check(bz_nspace_cclip(as_ptr<const double>(G), nt, n_real, as_ptr<const double>(a0), c_tau, M,
                          eps, as_ptr<float>(out), as_stream(stream)),
          "nspace_cclip");
  });

  m.def("nspace_preagg", [](uint64_t G, int n, int mode, double param, int iparam, uint64_t W, uint64_t W32,
                            uint64_t stream) {
    check(bz_nspace_preagg(as_ptr<const double>(G), n, mode, param, iparam, as_ptr<double>(W), as_ptr<float>(W32),
                           as_stream(stream)),
          "nspace_preagg");
  });
  m.def("nspace_caf", [](uint64_t G, int n, int f, int power_iters, uint64_t out, uint64_t stream) {
    check(bz_nspace_caf(as_ptr<const double>(G), n, f, power_iters, as_ptr<float>(out), as_stream(stream)),
          "nspace_caf");
  });

  bz_bind_runtime(m);
  bz_bind_vmm(m);
}
