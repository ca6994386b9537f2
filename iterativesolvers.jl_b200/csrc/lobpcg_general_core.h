// lobpcg_general_core.h -- lobpcg(A, B, largest, X0; P, C, tol, maxiter) of reference src/lobpcg.jl:827-893 with the
// step functor :692-749 in its GENERAL form: the generalized problem A x = lambda B x (B != nothing: B-blocks, B-inner
// products in CholQR :365-393 and in the Gram matrices :262-338), operators and preconditioner given as callbacks, and
// the constraint in the B inner product (:144-224).  Written as fused passes (pass_core.h) over column-major n x bs
// blocks: correct and complete first.  The standard problem on a b200_csr keeps the tuned engine of lobpcg.cu (row-major
// blocks, TMA-fed Gram kernels, tensor-pipe update); this engine re-reads a block once per column of the other factor
// of a Gram product (bs x (1 + bs) column reads instead of 2 bs) and applies the operators column by column.
//
// Block primitives (all reuse the constraint passes of lobpcg_constraint_core.h):
//   gram(L, R)        G[k][j] = <L_k, R_j>            one pass per column of L (16 sums)            :262-271
//   combine(Out, ...) Out = sum_b In_b M_b             zero + one pass per 16 columns of every In_b   :645-689
//   rdiv(X, U)        X <- X U^-1 (U upper)            one pass, the column sweep of rdiv! per row    :345-355
//   residual          R = AX - BX diag(lambda), ||.||  one pass (16 sums)                             :533-547
#pragma once
#include <algorithm>
#include <vector>

#include "dense_small.h"
#include "lobpcg_constraint_core.h"

namespace b200 {

// ---- rdiv!(X, UpperTriangular(U)) :345-355 on the rows of a column-major block: x_j = (x_j - sum_{i<j} x_i U[i,j]) / U[j,j]
template <typename T>
struct BlkRdiv {
  static constexpr int NRED = 0;
  T *X;
  int64_t ld;
  int bs;
  T U[kConBlock][kConBlock];          // U[i][j], upper triangle used
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *) const {
    T x[kConBlock];
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j) x[j] = j < bs ? X[i + j * ld] : (T)0;
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) {
        T v = x[j];
        B200_UNROLL
        for (int c = 0; c < kConBlock; ++c)
          if (c < j) v = v - x[c] * U[c][j];
        x[j] = v / U[j][j];
      }
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) X[i + j * ld] = x[j];
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

// ---- residuals! :533-547: R = AX - BX * Diagonal(lambda) ; sums of squares per column
template <typename T>
struct BlkResidual {
  static constexpr int NRED = kConBlock;
  const T *AX, *BX;
  T *R;
  int64_t ld;
  int bs;
  double *out;                        // kConBlock doubles (device): sums, also the allreduce buffer
  T lam[kConBlock];
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *acc) const {
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) {
        const T r = AX[i + j * ld] - BX[i + j * ld] * lam[j];
        R[i + j * ld] = r;
        acc[j] += (double)r * (double)r;
      }
  }
  B200_HD double *sums() const { return out; }
  B200_HD void finish(const double *tot) const {
    for (int j = 0; j < kConBlock; ++j) out[j] = tot[j];
  }
};

// ---- RPreconditioner with a diagonal M :236-242
template <typename T>
struct BlkJacobi {
  static constexpr int NRED = 0;
  T *X;
  int64_t ld;
  int bs;
  const T *d;
  B200_HD bool skip() const { return false; }
  B200_HD void load() {}
  B200_HD void elem(int64_t i, double *) const {
    const T di = d[i];
    B200_UNROLL
    for (int j = 0; j < kConBlock; ++j)
      if (j < bs) X[i + j * ld] = X[i + j * ld] / di;
  }
  B200_HD double *sums() const { return nullptr; }
  B200_HD void finish(const double *) const {}
};

struct LobpcgGenOutcome {
  int64_t iterations;
  int converged, status;              // status: 0, 1 = PosDefException in CholQR (:380), 2 = in the Rayleigh-Ritz problem
};

template <typename T, typename B>
struct LobpcgGen {
  B &be;
  int64_t n, ld;
  double *g_dev;                      // 16 x 16 doubles
  std::vector<double> g_host;

  // G (bl x 16 row-major, host) = L' R
  int gram(const T *L, int bl, const T *R, int br, double *G) {
    int st = constraint_gram<T>(be, L, ld, bl, R, 1, ld, br, n, g_dev, g_host.data());
    if (st) return st;
    for (int k = 0; k < bl; ++k)
      for (int j = 0; j < kConBlock; ++j) G[k * kConBlock + j] = g_host[(size_t)k * kConBlock + j];
    return 0;
  }
  // Out (l columns) = sum_b In_b (kk_b columns) * M_b (kk_b x l, column-major with leading dimension ldm_b)
  struct Term {
    const T *In;
    int kk;
    const double *M;
    int ldm;
  };
  int combine(T *Out, int l, const std::vector<Term> &terms) {
    int st;
    for (int j = 0; j < l; ++j)
      if ((st = be.zero(Out + (int64_t)j * ld, sizeof(T) * (size_t)n))) return st;
    for (const Term &t : terms) {
      if (t.kk == 0) continue;
      std::vector<double> coef((size_t)t.kk * kConBlock, 0.0);
      for (int c = 0; c < t.kk; ++c)
        for (int j = 0; j < l; ++j) coef[(size_t)c * kConBlock + j] = -t.M[c + (size_t)j * t.ldm];
      if ((st = constraint_update<T>(be, t.In, ld, t.kk, coef.data(), Out, 1, ld, l, n))) return st;
    }
    return 0;
  }
  int copy_block(T *Out, const T *In, int l) {
    int st;
    for (int j = 0; j < l; ++j)
      if ((st = be.copy(Out + (int64_t)j * ld, In + (int64_t)j * ld, sizeof(T) * (size_t)n))) return st;
    return 0;
  }
  int apply_block(const typename B::Op *Op, const T *In, T *Out, int l) {
    int st;
    for (int j = 0; j < l; ++j)
      if ((st = be.apply(Op, In + (int64_t)j * ld, Out + (int64_t)j * ld))) return st;
    return 0;
  }
  int rdiv(T *X, const double *U, int bs) {     // U: bs x bs column-major upper factor
    BlkRdiv<T> f;
    f.X = X;
    f.ld = ld;
    f.bs = bs;
    for (int i = 0; i < kConBlock; ++i)
      for (int j = 0; j < kConBlock; ++j) f.U[i][j] = (i < bs && j < bs && i <= j) ? (T)U[i + (size_t)j * bs] : (T)(i == j);
    return be.pass(f, n);
  }
  // CholQR :365-393: X (and AX, BX when given) <- . * R^-1 with R'R = X' BX.  Returns 1 on PosDefException.
  int cholqr(T *X, T *BX /* == X when not generalized */, T *AX, int bs, bool generalized, int *posdef) {
    double G[kConBlock * kConBlock];
    int st = gram(X, bs, BX, bs, G);
    if (st) return st;
    std::vector<double> U((size_t)bs * bs);
    for (int i = 0; i < bs; ++i)
      for (int j = 0; j < bs; ++j) U[i + (size_t)j * bs] = i <= j ? G[i * kConBlock + j] : G[j * kConBlock + i];   // Hermitian(gram): upper
    for (int i = 0; i < bs; ++i) U[i + (size_t)i * bs] = G[i * kConBlock + i];                                    // realdiag! :373
    if (dense::cholesky_upper(U.data(), bs, bs)) {
      *posdef = 1;
      return 0;
    }
    if ((st = rdiv(X, U.data(), bs))) return st;                                     // :384
    if (AX && (st = rdiv(AX, U.data(), bs))) return st;                              // update_AX :385
    if (generalized && BX != X && (st = rdiv(BX, U.data(), bs))) return st;          // update_BX :386
    return 0;
  }
};

// A, Bop (nullptr: standard problem), Pop (callback preconditioner or nullptr), jac (Jacobi diagonal or nullptr);
// constraint: Y, BY (n x nc, ldy; BY == Y for the standard problem), U (host upper factor of Y'BY), nc (0: none).
// X: n x sizeX column-major with leading dimension ldx (overwritten by the Ritz vectors).
template <typename T, typename B>
int lobpcg_general_run(B &be, const typename B::Op *A, const typename B::Op *Bop, const typename B::Op *Pop, const T *jac,
                       const T *Y, const T *BY, int64_t ldy, int nc, const double *Ucon, T *Xuser, int64_t ldx,
                       int sizeX, int64_t n, int largest, double tol, int64_t maxiter, int fixed_iterations,
                       double *lambda_host, double *resnorm_host, LobpcgGenOutcome *out, double *trace_resnorm = nullptr,
                       double *trace_ritz = nullptr, int64_t trace_cap = 0) {
  const bool generalized = Bop != nullptr;
  if (tol < 0) tol = pow(eps_of<T>(), 0.3);                                          // default_tolerance :751
  if (maxiter < 0) maxiter = 200;                                                    // :865
  const int64_t ld = (int64_t)((((sizeof(T) * (size_t)(n > 0 ? n : 1)) + 255) / 256 * 256) / sizeof(T));
  const size_t blk = sizeof(T) * (size_t)ld * sizeX;
  const int nblocks = 17;
  void *ws = nullptr;
  int st = be.workspace(blk * nblocks + sizeof(double) * (size_t)(kConBlock * kConBlock + nc * kConBlock + 64), &ws);
  if (st) return st;
  char *w = (char *)ws;
  auto take = [&]() { T *p = (T *)w; w += blk; return p; };
  T *X = take(), *AX = take(), *BXs = take(), *R = take(), *P = take(), *AP = take(), *BPs = take();
  T *aR = take(), *aAR = take(), *aBRs = take(), *aP = take(), *aAP = take(), *aBPs = take();
  T *T1 = take(), *T2 = take(), *T3 = take(), *PT = take();
  double *g_dev = (double *)w; w += sizeof(double) * kConBlock * kConBlock;
  double *gc_dev = (double *)w;                                                      // constraint Gram scratch: nc x 16
  T *BX = generalized ? BXs : X;
  T *BP = generalized ? BPs : P;
  LobpcgGen<T, B> L{be, n, ld, g_dev, std::vector<double>((size_t)kConBlock * kConBlock)};
  std::vector<double> gc_host((size_t)std::max(nc, 1) * kConBlock);
  auto constrain = [&](T *Blk, int bs) -> int {                                      // (constr!)(X, temp) :212-224
    if (nc <= 0) return 0;
    return constraint_apply<T>(be, Y, ldy, nc, Ucon, Blk, 1, ld, bs, n, gc_dev, gc_host.data()) ? -1 : 0;
  };
  // NOTE: with B != I the reference projects with BY' X (:217); constraint_apply takes the basis whose inner products
  // are formed (BY) and the one that is subtracted (Y) as one array when they alias -- here they may differ:
  auto constrain_b = [&](T *Blk, int bs) -> int {
    if (nc <= 0) return 0;
    if (BY == Y) return constrain(Blk, bs);
    int s2 = constraint_gram<T>(be, BY, ldy, nc, Blk, 1, ld, bs, n, gc_dev, gc_host.data());
    if (s2) return s2;
    con_chol_solve(Ucon, nc, gc_host.data(), bs);
    return constraint_update<T>(be, Y, ldy, nc, gc_host.data(), Blk, 1, ld, bs, n);
  };

  for (int j = 0; j < sizeX; ++j)
    if ((st = be.copy(X + (int64_t)j * ld, Xuser + (int64_t)j * ldx, sizeof(T) * (size_t)n))) return st;
  if ((st = constrain_b(X, sizeX))) return st;                                       // iterator.constr!(X, temp) :868 / :875

  std::vector<double> ritz(3 * sizeX, 0.0), residuals(sizeX, NAN);                    // :473-477
  std::vector<char> mask(sizeX, 1);
  int bs = sizeX, status = 0, posdef = 0;
  int64_t iteration = 1;

  auto select = [&](const std::vector<double> &wv, int sub, std::vector<int> &perm) {  // partialsortperm!(...; rev=largest) :623
    perm.resize(sub);
    for (int i = 0; i < sub; ++i) perm[i] = i;
    if (largest) std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return wv[a] > wv[b]; });
    for (int j = 0; j < sizeX; ++j) ritz[j] = wv[perm[j]];
  };
  auto residuals_ = [&]() -> int {                                                    // residuals! :533-547
    BlkResidual<T> f;
    f.AX = AX; f.BX = BX; f.R = R; f.ld = ld; f.bs = sizeX; f.out = g_dev;
    for (int j = 0; j < kConBlock; ++j) f.lam[j] = j < sizeX ? (T)ritz[j] : (T)0;
    int s2 = be.pass(f, n);
    if (s2) return s2;
    double nn[kConBlock];
    if ((s2 = be.to_host(nn, g_dev, sizeof(nn)))) return s2;
    for (int j = 0; j < sizeX; ++j) residuals[j] = sqrt(nn[j]);
    return 0;
  };

  while (iteration <= maxiter) {                                                      // :880
    if (iteration == 1) {                                                             // :695-703
      if (generalized && (st = L.apply_block(Bop, X, BX, sizeX))) return st;          // B_mul_X! :526
      if ((st = L.cholqr(X, BX, nullptr, sizeX, generalized, &posdef))) return st;    // ortho! :528
      if (posdef) { status = 1; break; }
      if ((st = L.apply_block(A, X, AX, sizeX))) return st;                           // A_mul_X! :530
      double G[kConBlock * kConBlock];
      if ((st = L.gram(X, sizeX, AX, sizeX, G))) return st;                           // XAX :262
      std::vector<double> Am((size_t)sizeX * sizeX), wv, Z;
      for (int i = 0; i < sizeX; ++i)
        for (int j = 0; j < sizeX; ++j) Am[i + (size_t)j * sizeX] = i <= j ? G[i * kConBlock + j] : G[j * kConBlock + i];
      if (dense::sym_eig(Am, sizeX, wv, Z)) { status = 2; break; }
      std::vector<int> perm;
      select(wv, sizeX, perm);
      std::vector<double> V((size_t)sizeX * sizeX);
      for (int j = 0; j < sizeX; ++j)
        for (int i = 0; i < sizeX; ++i) V[i + (size_t)j * sizeX] = Z[i + (size_t)perm[j] * sizeX];
      // update_X_P!(0, 0) :629-690: X = X V ; AX = AX V ; BX = BX V
      if ((st = L.combine(T1, sizeX, {{X, sizeX, V.data(), sizeX}}))) return st;
      if ((st = L.combine(T2, sizeX, {{AX, sizeX, V.data(), sizeX}}))) return st;
      if (generalized) {
        if ((st = L.combine(T3, sizeX, {{BX, sizeX, V.data(), sizeX}}))) return st;
        if ((st = L.copy_block(BX, T3, sizeX))) return st;
      }
      if ((st = L.copy_block(X, T1, sizeX))) return st;
      if ((st = L.copy_block(AX, T2, sizeX))) return st;
      if ((st = residuals_())) return st;
    } else {
      const bool with_p = iteration > 2;
      int idx[kConBlock], k = 0;                                                      // update_active! :557-562
      for (int j = 0; j < sizeX; ++j)
        if (mask[j]) idx[k++] = j;
      for (int c = 0; c < bs; ++c) {
        if ((st = be.copy(aR + (int64_t)c * ld, R + (int64_t)idx[c] * ld, sizeof(T) * (size_t)n))) return st;
        if (with_p) {
          if ((st = be.copy(aP + (int64_t)c * ld, P + (int64_t)idx[c] * ld, sizeof(T) * (size_t)n))) return st;
          if ((st = be.copy(aAP + (int64_t)c * ld, AP + (int64_t)idx[c] * ld, sizeof(T) * (size_t)n))) return st;
          if (generalized && (st = be.copy(aBPs + (int64_t)c * ld, BP + (int64_t)idx[c] * ld, sizeof(T) * (size_t)n))) return st;
        }
      }
      T *aBR = generalized ? aBRs : aR;
      T *aBP = generalized ? aBPs : aP;
      // precond_constr! :564-569
      if (Pop) {                                                                      // ldiv!(buffer, M, X); X .= buffer :238-240
        if ((st = L.apply_block(Pop, aR, PT, bs))) return st;
        if ((st = L.copy_block(aR, PT, bs))) return st;
      } else if (jac) {
        if ((st = be.pass(BlkJacobi<T>{aR, ld, bs, jac}, n))) return st;
      }
      if ((st = constrain_b(aR, bs))) return st;                                      // :567
      if (generalized && (st = L.apply_block(Bop, aR, aBR, bs))) return st;           // ortho_AB_mul_X! :524-532
      if ((st = L.cholqr(aR, aBR, nullptr, bs, generalized, &posdef))) return st;
      if (posdef) { status = 1; break; }
      if ((st = L.apply_block(A, aR, aAR, bs))) return st;
      if (with_p) {
        if ((st = L.cholqr(aP, aBP, aAP, bs, generalized, &posdef))) return st;       // :733
        if (posdef) { status = 1; break; }
      }
      const int n1 = sizeX, n2 = bs, n3 = with_p ? bs : 0, sub = n1 + n2 + n3;
      std::vector<double> gA((size_t)sub * sub, 0.0), gB((size_t)sub * sub, 0.0);
      auto setA = [&](int i, int j, double v) { gA[i + (size_t)j * sub] = v; gA[j + (size_t)i * sub] = v; };
      auto setB = [&](int i, int j, double v) { gB[i + (size_t)j * sub] = v; gB[j + (size_t)i * sub] = v; };
      for (int i = 0; i < n1; ++i) setA(i, i, ritz[i]);                               // Diagonal(lambda) :289
      for (int i = 0; i < sub; ++i) setB(i, i, 1.0);                                  // I! :315,322,331
      double G[kConBlock * kConBlock];
      if ((st = L.gram(X, n1, aAR, n2, G))) return st;                                // XAR :265
      for (int i = 0; i < n1; ++i) for (int j = 0; j < n2; ++j) setA(i, n1 + j, G[i * kConBlock + j]);
      if ((st = L.gram(X, n1, aBR, n2, G))) return st;                                // XBR :270
      for (int i = 0; i < n1; ++i) for (int j = 0; j < n2; ++j) setB(i, n1 + j, G[i * kConBlock + j]);
      if ((st = L.gram(aR, n2, aAR, n2, G))) return st;                               // RAR :266
      for (int i = 0; i < n2; ++i) for (int j = i; j < n2; ++j) setA(n1 + i, n1 + j, G[i * kConBlock + j]);
      if (with_p) {
        if ((st = L.gram(X, n1, aAP, n3, G))) return st;                              // XAP :264
        for (int i = 0; i < n1; ++i) for (int j = 0; j < n3; ++j) setA(i, n1 + n2 + j, G[i * kConBlock + j]);
        if ((st = L.gram(X, n1, aBP, n3, G))) return st;                              // XBP :269
        for (int i = 0; i < n1; ++i) for (int j = 0; j < n3; ++j) setB(i, n1 + n2 + j, G[i * kConBlock + j]);
        if ((st = L.gram(aAR, n2, aP, n3, G))) return st;                             // RAP :267
        for (int i = 0; i < n2; ++i) for (int j = 0; j < n3; ++j) setA(n1 + i, n1 + n2 + j, G[i * kConBlock + j]);
        if ((st = L.gram(aBR, n2, aP, n3, G))) return st;                             // RBP :271
        for (int i = 0; i < n2; ++i) for (int j = 0; j < n3; ++j) setB(n1 + i, n1 + n2 + j, G[i * kConBlock + j]);
        if ((st = L.gram(aP, n3, aAP, n3, G))) return st;                             // PAP :268
        for (int i = 0; i < n3; ++i) for (int j = i; j < n3; ++j) setA(n1 + n2 + i, n1 + n2 + j, G[i * kConBlock + j]);
      }
      std::vector<double> wv, Z;
      if (dense::sym_eig_generalized(gA, gB, sub, wv, Z)) { status = 2; break; }      // :622
      std::vector<int> perm;
      select(wv, sub, perm);
      std::vector<double> Vx((size_t)n1 * sizeX), Vr((size_t)std::max(n2, 1) * sizeX), Vp((size_t)std::max(n3, 1) * sizeX);
      for (int j = 0; j < sizeX; ++j) {
        const double *z = &Z[(size_t)perm[j] * sub];
        for (int i = 0; i < n1; ++i) Vx[i + (size_t)j * n1] = z[i];
        for (int i = 0; i < n2; ++i) Vr[i + (size_t)j * n2] = z[n1 + i];
        for (int i = 0; i < n3; ++i) Vp[i + (size_t)j * n3] = z[n1 + n2 + i];
      }
      // update_X_P! :645-689: P = aR Vr + aP Vp (and A-, B- twins) ; X = X Vx + P ...
      if ((st = L.combine(T1, sizeX, {{aR, n2, Vr.data(), n2}, {aP, n3, Vp.data(), n3}}))) return st;
      if ((st = L.combine(T2, sizeX, {{aAR, n2, Vr.data(), n2}, {aAP, n3, Vp.data(), n3}}))) return st;
      if (generalized) {
        if ((st = L.combine(T3, sizeX, {{aBR, n2, Vr.data(), n2}, {aBP, n3, Vp.data(), n3}}))) return st;
        if ((st = L.copy_block(BPs, T3, sizeX))) return st;
      }
      if ((st = L.copy_block(P, T1, sizeX))) return st;
      if ((st = L.copy_block(AP, T2, sizeX))) return st;
      std::vector<double> Id((size_t)sizeX * sizeX, 0.0);
      for (int i = 0; i < sizeX; ++i) Id[i + (size_t)i * sizeX] = 1.0;
      if ((st = L.combine(T1, sizeX, {{X, n1, Vx.data(), n1}, {P, sizeX, Id.data(), sizeX}}))) return st;
      if ((st = L.combine(T2, sizeX, {{AX, n1, Vx.data(), n1}, {AP, sizeX, Id.data(), sizeX}}))) return st;
      if (generalized) {
        if ((st = L.combine(T3, sizeX, {{BX, n1, Vx.data(), n1}, {BPs, sizeX, Id.data(), sizeX}}))) return st;
        if ((st = L.copy_block(BX, T3, sizeX))) return st;
      }
      if ((st = L.copy_block(X, T1, sizeX))) return st;
      if ((st = L.copy_block(AX, T2, sizeX))) return st;
      if ((st = residuals_())) return st;
    }
    if (iteration - 1 < trace_cap) {                               // log = true: LOBPCGState(iteration, residuals, ritz_values) :744-745
      for (int j = 0; j < sizeX; ++j) {
        if (trace_resnorm) trace_resnorm[(iteration - 1) * sizeX + j] = residuals[j];
        if (trace_ritz) trace_ritz[(iteration - 1) * sizeX + j] = ritz[j];
      }
    }
    bs = 0;                                                                           // update_mask! :549-555
    for (int j = 0; j < sizeX; ++j) {
      mask[j] = fixed_iterations ? 1 : (residuals[j] > tol);
      bs += mask[j];
    }
    if (bs == 0) break;                                                               // :885
    iteration += 1;                                                                   // :886
  }
  for (int j = 0; j < sizeX; ++j)
    if ((st = be.copy(Xuser + (int64_t)j * ldx, X + (int64_t)j * ld, sizeof(T) * (size_t)n))) return st;
  int sync = 0;
  if ((st = be.read_flag((const int *)g_dev, &sync))) return st;                      // completes the copies before returning
  bool conv = true;
  for (int j = 0; j < sizeX; ++j) {
    if (lambda_host) lambda_host[j] = ritz[j];
    if (resnorm_host) resnorm_host[j] = residuals[j];
    conv = conv && (residuals[j] <= tol);
  }
  out->iterations = iteration;                                                        // :890
  out->converged = conv;
  out->status = status;
  return 0;
}

}  // namespace b200
