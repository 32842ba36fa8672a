// myo_solver.cuh -- constraint assembly, primal Newton solver and semi-implicit Euler (one env per warp).
// Semantics: MuJoCo's soft-constraint model: rows = joint equalities, joint
// limits, pyramidal/frictionless contacts; cost 1/2 (a-a0)'M(a-a0) + sum_i s_i(J_i a - aref_i);
// Newton with exact line search on H = M + J' D_active J (register/shuffle Cholesky of the dense packed H; tree-sparse L'DL when no contact is active).
#pragma once
#include "myo_device.cuh"

struct Solv {   // scratch views (stage 3/4)
  double *H, *con, *conJ, *D, *aref, *jar, *jv, *a, *g, *p, *Ma, *Mp, *eqJ, *Hs, *LD, *Dinv;
  int *cpair, *crow, *cnrow, *lrow;   // contact pair idx, first efc row, #rows ; limit row descriptors (dof | side << 16)
};
__device__ __forceinline__ Solv solv_views(const DevModel& m, const Warp& w) {
  Solv s; double* A = w.scr;
  s.H = A + m.s_H; s.con = A + m.s_con; s.conJ = A + m.s_conJ; s.D = A + m.s_efD; s.aref = A + m.s_efA; s.jar = A + m.s_efR; s.jv = A + m.s_efV;
  s.a = A + m.s_va; s.g = A + m.s_vg; s.p = A + m.s_vp; s.Ma = A + m.s_vMa; s.Mp = A + m.s_vMp; s.eqJ = A + m.s_eqJ; s.Hs = A + m.s_Hs; s.LD = A + m.s_LD; s.Dinv = A + m.s_Dinv;
  int* ic = (int*)(A + m.s_icon); s.cpair = ic; s.crow = ic + m.maxcon; s.cnrow = ic + 2*m.maxcon; s.lrow = ic + 3*m.maxcon;
  return s; }

__device__ __forceinline__ double impedance(const double* si, double pos, double margin) {
  if (si[0] == si[1] || si[2] <= MYO_MINVAL) return 0.5*(si[0]+si[1]);
  double x = fabs((pos-margin)/si[2]);
  if (x >= 1 || x <= 0) return x >= 1 ? si[1] : si[0];
  double y;
  if (si[4] == 1) y = x;
  else if (si[4] == 2) y = x <= si[3] ? x*x/si[3] : 1-(1-x)*(1-x)/(1-si[3]);   // the default power, without pow()
  else if (x <= si[3]) y = pow(x, si[4])/pow(si[3], si[4]-1);
  else y = 1-pow(1-x, si[4])/pow(1-si[3], si[4]-1);
  return si[0]+y*(si[1]-si[0]); }

// y = M x using the per-row non-zero lists
__device__ __forceinline__ void mul_M(const DevModel& m, const Warp& w, double* y, const double* x) {
  const idx_t* radr = CI(PROW_adr); const idx_t* rcol = CI(PROW_col); const idx_t* ridx = CI(PROW_idx);
  for (int i = w.lane; i < m.nv; i += 32) { double s = 0; for (int e = radr[i]; e < radr[i+1]; e++) s += w.qM[ridx[e]]*x[rcol[e]]; y[i] = s; } }

// out[r] = (J x)_r for every constraint row
__device__ void rows_apply(const DevModel& m, const Warp& w, const Solv& s, const double* x, double* out) {
  const idx_t* eq = CI(PEQ);
  for (int e = w.lane; e < m.neq; e += 32) { double v = x[eq[PEQ_ISTRIDE*e+1]]; if (eq[PEQ_ISTRIDE*e+3] >= 0) v += s.eqJ[e]*x[eq[PEQ_ISTRIDE*e+3]]; out[e] = v; }
  for (int r = w.lane; r < w.nlimrow; r += 32) { int dsc = s.lrow[r]; double sg = (dsc >> 16) & 1 ? -1.0 : 1.0; out[m.neq + r] = sg*x[dsc & 0xffff]; }
  const idx_t* pr = CI(PPAIR); const double* pd = CD(PPAIR_d); const idx_t* path = CI(PPATH);
  for (int c = w.lane; c < w.ncon; c += 32) { int nr = s.cnrow[c]; if (!nr) continue;
    const idx_t* q = pr + PPAIR_ISTRIDE*s.cpair[c]; const double* J = s.conJ + (size_t)c*3*m.maxpath; double n = 0, t1 = 0, t2 = 0;
    for (int e = 0; e < q[4]; e++) { double xv = x[path[q[3]+e] >> 1]; n += J[3*e]*xv; t1 += J[3*e+1]*xv; t2 += J[3*e+2]*xv; }
    int rb = s.crow[c];
    if (nr == 1) out[rb] = n;
    else { const double* P = pd + q[6]*PPAIR_STRIDE; double mu1 = P[2], mu2 = P[3]; out[rb] = n+mu1*t1; out[rb+1] = n-mu1*t1; out[rb+2] = n+mu2*t2; out[rb+3] = n-mu2*t2; } }
}

// vec[d] += sum_r J[r][d] * wgt[r]  (wgt already includes D and the active mask)
__device__ void rows_applyT_add(const DevModel& m, const Warp& w, const Solv& s, const double* wgt, double* vec) {
  const idx_t* eq = CI(PEQ);
  if (w.lane == 0) for (int e = 0; e < m.neq; e++) { vec[eq[PEQ_ISTRIDE*e+1]] += wgt[e]; if (eq[PEQ_ISTRIDE*e+3] >= 0) vec[eq[PEQ_ISTRIDE*e+3]] += s.eqJ[e]*wgt[e]; }
  __syncwarp();
  for (int pass = 0; pass < 2; pass++) {
    for (int r = w.lane; r < w.nlimrow; r += 32) { int dsc = s.lrow[r]; int neg = (dsc >> 16) & 1; if (neg == pass) vec[dsc & 0xffff] += (neg ? -1.0 : 1.0)*wgt[m.neq + r]; }
    __syncwarp(); }
  const idx_t* pr = CI(PPAIR); const double* pd = CD(PPAIR_d); const idx_t* path = CI(PPATH);
  for (int c = 0; c < w.ncon; c++) { int nr = s.cnrow[c]; if (!nr) continue;
    const idx_t* q = pr + PPAIR_ISTRIDE*s.cpair[c]; const double* J = s.conJ + (size_t)c*3*m.maxpath; int rb = s.crow[c]; double wn, w1 = 0, w2 = 0;
    if (nr == 1) wn = wgt[rb];
    else { const double* P = pd + q[6]*PPAIR_STRIDE; wn = wgt[rb]+wgt[rb+1]+wgt[rb+2]+wgt[rb+3]; w1 = P[2]*(wgt[rb]-wgt[rb+1]); w2 = P[3]*(wgt[rb+2]-wgt[rb+3]); }
    for (int e = w.lane; e < q[4]; e += 32) vec[path[q[3]+e] >> 1] += J[3*e]*wn + J[3*e+1]*w1 + J[3*e+2]*w2;
    __syncwarp(); }
}

// ------------------------------------------------------------------ constraint assembly
__device__ void phase_constraints(const DevModel& m, Warp& w) {
  Solv s = solv_views(m, w);
  // joint equalities (always active)
  const idx_t* eq = CI(PEQ); const double* eqd = CD(PEQ_d);
  for (int e = w.lane; e < m.neq; e += 32) { const double* c = eqd + e*PEQ_STRIDE; int q1 = eq[PEQ_ISTRIDE*e], d1 = eq[PEQ_ISTRIDE*e+1], q2 = eq[PEQ_ISTRIDE*e+2], d2 = eq[PEQ_ISTRIDE*e+3];
    double pos0 = w.qpos[q1]-c[5], cpos, deriv = 0, vel = w.qvel[d1];
    if (q2 >= 0) { double x = w.qpos[q2]-c[6]; cpos = pos0-(c[0]+x*(c[1]+x*(c[2]+x*(c[3]+x*c[4])))); deriv = c[1]+x*(2*c[2]+x*(3*c[3]+x*4*c[4])); vel -= deriv*w.qvel[d2]; }
    else cpos = pos0-c[0];
    s.eqJ[e] = -deriv;
    double imp = impedance(c+10, cpos, 0), R = fmax(MYO_MINVAL, (1-imp)*c[7]/imp);
    s.D[e] = 1.0/R; s.aref[e] = -c[9]*vel - c[8]*imp*cpos; }
  // joint limits (one-sided)
  const idx_t* lim = CI(PLIM); const double* limd = CD(PLIM_d); int nrow = 0;
  for (int base = 0; base < m.nlim; base += 32) { int l = base + w.lane; bool lo = false, hi = false; double dlo = 0, dhi = 0; const double* c = limd + (l < m.nlim ? l : 0)*PLIM_STRIDE; int d = 0;
    if (l < m.nlim) { d = lim[2*l]; double q = w.qpos[lim[2*l+1]]; dlo = q-c[0]; dhi = c[1]-q; lo = dlo < c[2]; hi = dhi < c[2]; }
    unsigned m0 = __ballot_sync(FULL, lo), m1 = __ballot_sync(FULL, hi), lt = (1u << w.lane)-1; int idx = nrow + __popc(m0 & lt) + __popc(m1 & lt);
    for (int side = 0; side < 2; side++) { if (!(side ? hi : lo)) continue;
      double dist = side ? dhi : dlo, sg = side ? -1.0 : 1.0; int r = m.neq + idx; idx++;
      s.lrow[r - m.neq] = d | (side << 16);
      double imp = impedance(c+6, dist, c[2]), R = fmax(MYO_MINVAL, (1-imp)*c[3]/imp);
      s.D[r] = 1.0/R; s.aref[r] = -c[5]*sg*w.qvel[d] - c[4]*imp*(dist-c[2]); }
    nrow += __popc(m0) + __popc(m1); }
  w.nlimrow = nrow;
  // contacts: Jacobian over the dofs between the two bodies, regulariser, reference acceleration
  const idx_t* pr = CI(PPAIR); const double* pd = CD(PPAIR_d); const idx_t* path = CI(PPATH);
  int rowbase = m.neq + nrow;
  for (int base = 0; base < w.ncon; base += 32) { int c = base + w.lane; int nr = 0;
    if (c < w.ncon) { const idx_t* q = pr + PPAIR_ISTRIDE*s.cpair[c]; const double* P = pd + q[6]*PPAIR_STRIDE; double dist = s.con[c*CON_STRIDE];
      nr = (dist < P[0]-P[1]) ? (q[2] == 1 ? 1 : 4) : 0; }
    int incl = nr;   // inclusive warp scan
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (w.lane >= o) incl += t; }
    int total = __shfl_sync(FULL, incl, 31);
    if (c < w.ncon) { s.crow[c] = rowbase + incl - nr; s.cnrow[c] = nr; }
    if (c < w.ncon && nr) { const idx_t* q = pr + PPAIR_ISTRIDE*s.cpair[c]; const double* P = pd + q[6]*PPAIR_STRIDE; const double* cd = s.con + c*CON_STRIDE;
      const double* pos = cd + 1; double f[9]; for (int k = 0; k < 6; k++) f[k] = cd[4+k]; cross3(f+6, f, f+3); double* J = s.conJ + (size_t)c*3*m.maxpath; double vn = 0, v1 = 0, v2 = 0;
      for (int e = 0; e < q[4]; e++) { int code = path[q[3]+e], d = code >> 1; double sg = (code & 1) ? 1.0 : -1.0, cv[3]; dof_point_vel(m, w, d, pos, cv);
        double jn = sg*dot3(f, cv), j1 = sg*dot3(f+3, cv), j2 = sg*dot3(f+6, cv); J[3*e] = jn; J[3*e+1] = j1; J[3*e+2] = j2;
        double qd = w.qvel[d]; vn += jn*qd; v1 += j1*qd; v2 += j2*qd; }
      double dist = cd[0], inc = P[0]-P[1], imp = impedance(P+7, dist, inc), K = P[5], B = P[6], tran = CD(PPAIR_tran)[s.cpair[c]]; int rb = rowbase + incl - nr;
      if (nr == 1) { double R = fmax(MYO_MINVAL, (1-imp)*tran/imp); s.D[rb] = 1.0/R; s.aref[rb] = -B*vn - K*imp*(dist-inc); }
      else { double mu1 = P[2], mu2 = P[3]; double R0 = fmax(MYO_MINVAL, (1-imp)*(tran+mu1*mu1*tran)/imp), Rpy = 2*mu1*mu1*R0, Dv = 1.0/Rpy, kp = K*imp*(dist-inc);
        s.D[rb] = s.D[rb+1] = s.D[rb+2] = s.D[rb+3] = Dv;
        s.aref[rb] = -B*(vn+mu1*v1)-kp; s.aref[rb+1] = -B*(vn-mu1*v1)-kp; s.aref[rb+2] = -B*(vn+mu2*v2)-kp; s.aref[rb+3] = -B*(vn-mu2*v2)-kp; } }
    rowbase += total; }
  w.nefc = rowbase;
  __syncwarp();
}

// ------------------------------------------------------------------ dense Cholesky (packed lower triangle, in place) + solve in shared memory
#define TRI(i, j) ((i)*((i)+1)/2 + (j))
// t -> (i >= j) with t = i(i+1)/2 + j
__device__ __forceinline__ void tri_index(int t, int& i, int& j) {
  i = __float2int_rd((sqrtf(8.0f*t + 1.0f) - 1.0f)*0.5f); if (((i+1)*(i+2) >> 1) <= t) i++; if (((i*(i+1)) >> 1) > t) i--;
  j = t - ((i*(i+1)) >> 1); }

// In-place right-looking Cholesky with the whole trailing update of a column spread flat over the lanes (a short rolled loop: the
// code stays in the instruction cache, unlike the unrolled register version).  On exit column k holds L[.][k] below the diagonal;
// the diagonal keeps the PIVOT (not its root): *dinv_lane = 1/sqrt(pivot of row `lane`) for lane < n, and, if fix_diag, H[k][k] = sqrt(pivot).
__device__ __noinline__ void chol_factor(double* H, int n, int lane, double* dinv_lane, bool fix_diag) {
  double inv_prev = 0, dinv = 1.0;
  #pragma unroll 1
  for (int k = 0; k < n; k++) {
    double inv = rsqrt(fmax(H[TRI(k,k)], MYO_MINVAL)), inv2 = inv*inv;       // every lane reads the pivot: no broadcast step
    if (lane == k) dinv = inv;
    if (k > 0) for (int i = k + lane; i < n; i += 32) H[TRI(i,k-1)] *= inv_prev;   // scale the previous column (its readers are done)
    const int m_ = n-1-k, cnt = (m_*(m_+1)) >> 1; const double* colk = H + k;       // H[TRI(i,k)] = H[i(i+1)/2 + k]
    #pragma unroll 2
    for (int t = lane; t < cnt; t += 32) { int a, b; tri_index(t, a, b); int i = k+1+a, j = k+1+b;
      H[TRI(i,j)] -= colk[(i*(i+1)) >> 1]*colk[(j*(j+1)) >> 1]*inv2; }
    __syncwarp();
    inv_prev = inv; }
  if (fix_diag) { for (int k = lane; k < n; k += 32) H[TRI(k,k)] = sqrt(fmax(H[TRI(k,k)], MYO_MINVAL)); __syncwarp(); }
  *dinv_lane = dinv;
}
// n <= 32: substitutions with one row per lane (rhs in a register, exchanged by shuffles; L read from shared memory)
__device__ __noinline__ void chol_smem32(double* H, int n, double* x, int lane) {
  double dinv; chol_factor(H, n, lane, &dinv, false);
  double b = lane < n ? x[lane] : 0.0; const double* row = H + ((lane*(lane+1)) >> 1);
  #pragma unroll 1
  for (int k = 0; k < n; k++) { double yk = __shfl_sync(FULL, b*dinv, k); if (lane == k) b = yk; else if (lane > k && lane < n) b = fma(-row[k], yk, b); }
  #pragma unroll 1
  for (int k = n-1; k >= 0; k--) { double xk = __shfl_sync(FULL, b*dinv, k); if (lane == k) b = xk; else if (lane < k) b = fma(-H[TRI(k,lane)], xk, b); }
  if (lane < n) x[lane] = b;
  __syncwarp();
}
// row-per-lane variant (any n): the n > 32 fallback (legs, nv = 34: 176 k vs 167 k env-steps/s with the flat-update version above)
__device__ void chol_factor_rows(double* H, int n, int lane) {
  for (int k = 0; k < n; k++) {
    double dkk = sqrt(fmax(H[TRI(k,k)], MYO_MINVAL));
    __syncwarp();
    if (lane == 0) H[TRI(k,k)] = dkk;
    double inv = 1.0/dkk;
    for (int i = k+1+lane; i < n; i += 32) H[TRI(i,k)] *= inv;
    __syncwarp();
    for (int i = k+1+lane; i < n; i += 32) { double lik = H[TRI(i,k)]; for (int j = k+1; j <= i; j++) H[TRI(i,j)] -= lik*H[TRI(j,k)]; }
    __syncwarp(); }
}
// x <- H^-1 x  (H holds the Cholesky factor)
__device__ void chol_solve(const double* H, int n, double* x, int lane) {
  for (int k = 0; k < n; k++) { double xk = x[k]/H[TRI(k,k)]; __syncwarp(); if (lane == 0) x[k] = xk;
    for (int i = k+1+lane; i < n; i += 32) x[i] -= H[TRI(i,k)]*xk; __syncwarp(); }
  for (int k = n-1; k >= 0; k--) { double xk = x[k]/H[TRI(k,k)]; __syncwarp(); if (lane == 0) x[k] = xk;
    for (int i = lane; i < k; i += 32) x[i] -= H[TRI(k,i)]*xk; __syncwarp(); }
}
// Register-resident Cholesky + solve for n <= NMAX <= 32: lane i keeps row i of the lower triangle in registers, columns are
// exchanged with warp shuffles (no shared-memory latency, no barriers).  H: dense n x n in shared memory (read only); x: rhs in / solution out.
template <int NMAX>
__device__ __forceinline__ void chol_reg(const double* H, int n, double* x, int lane) {
  // The matrix is padded with an identity block up to NMAX so that every shuffle below sits in straight-line, branch-free code
  // (a shuffle under an `if (k < n)` makes the compiler wrap each one in a WARPSYNC/ENDCOLLECTIVE sequence: 6x the instructions).
  double r[NMAX], c[NMAX];      // r: row `lane` of L (lower part; the upper part holds unread garbage);  c: column `lane` of L, captured from the shuffles
  #pragma unroll
  for (int j = 0; j < NMAX; j++) { r[j] = (lane < n && j <= lane) ? H[TRI(lane, j)] : ((j == lane && lane >= n) ? 1.0 : 0.0); c[j] = 0.0; }
  double b = lane < n ? x[lane] : 0.0, dinv = 1.0;
  #pragma unroll
  for (int k = 0; k < NMAX; k++) {
    double inv = rsqrt(fmax(__shfl_sync(FULL, r[k], k), MYO_MINVAL));      // 1/sqrt(pivot)
    if (lane == k) dinv = inv;
    r[k] *= inv;                                   // lanes > k: L[i][k]; lane k: sqrt(pivot) (unused below); lanes < k: 0
    #pragma unroll
    for (int j = k+1; j < NMAX; j++) { double ljk = __shfl_sync(FULL, r[k], j); if (lane == k) c[j] = ljk; r[j] = fma(-r[k], ljk, r[j]); } }
  // forward substitution  L y = b
  #pragma unroll
  for (int k = 0; k < NMAX; k++) { double yk = __shfl_sync(FULL, b*dinv, k); b = lane == k ? yk : (lane > k ? fma(-r[k], yk, b) : b); }
  // backward substitution L' x = y: x_k is final once all j > k are eliminated; lane i < k holds L[k][i] in c[k]  (c[k] = 0 on lanes >= k)
  #pragma unroll
  for (int k = NMAX-1; k >= 0; k--) { double xk = __shfl_sync(FULL, b*dinv, k); b = lane == k ? xk : fma(-c[k], xk, b); }
  if (lane < n) x[lane] = b;
  __syncwarp();
}
__device__ __noinline__ void chol_reg8(const double* H, int n, double* x, int lane) { chol_reg<8>(H, n, x, lane); }
__device__ __noinline__ void chol_reg16(const double* H, int n, double* x, int lane) { chol_reg<16>(H, n, x, lane); }
__device__ __noinline__ void chol_reg24(const double* H, int n, double* x, int lane) { chol_reg<24>(H, n, x, lane); }
__device__ __noinline__ void chol_reg32(const double* H, int n, double* x, int lane) { chol_reg<32>(H, n, x, lane); }

// Bordered register Cholesky for 32 < n <= 32+E: H = [A B'; B C] with A the leading 32x32 block.  A = L11 L11' is factored in
// registers exactly like chol_reg<32>; the rows of B ride along as extra right-hand sides of the forward substitution
// (L21 = B L11^-T), the E x E Schur complement C - L21 L21' is reduced with warp sums and factored redundantly by every lane.
template <int E>
__device__ __noinline__ void chol_reg32b(const double* H, int n, double* x, int lane) {
  const int e = n - 32;
  double r[32], c[32], bq[E];
  #pragma unroll
  for (int j = 0; j < 32; j++) { r[j] = j <= lane ? H[TRI(lane, j)] : 0.0; c[j] = 0.0; }
  #pragma unroll
  for (int q = 0; q < E; q++) bq[q] = q < e ? H[TRI(32+q, lane)] : 0.0;
  double b = x[lane], dinv = 1.0;
  #pragma unroll
  for (int k = 0; k < 32; k++) {
    double inv = rsqrt(fmax(__shfl_sync(FULL, r[k], k), MYO_MINVAL));
    if (lane == k) dinv = inv;
    r[k] *= inv;
    #pragma unroll
    for (int j = k+1; j < 32; j++) { double ljk = __shfl_sync(FULL, r[k], j); if (lane == k) c[j] = ljk; r[j] = fma(-r[k], ljk, r[j]); } }
  #pragma unroll
  for (int k = 0; k < 32; k++) { double yk = __shfl_sync(FULL, b*dinv, k); b = lane == k ? yk : (lane > k ? fma(-r[k], yk, b) : b);
    #pragma unroll
    for (int q = 0; q < E; q++) { double yq = __shfl_sync(FULL, bq[q]*dinv, k); bq[q] = lane == k ? yq : (lane > k ? fma(-r[k], yq, bq[q]) : bq[q]); } }
  // b = y1[lane], bq[q] = L21[q][lane]; Schur complement (identity-padded beyond e) and its right-hand side
  double S[E][E], y2[E];
  #pragma unroll
  for (int q = 0; q < E; q++) {
    #pragma unroll
    for (int p = 0; p <= q; p++) { double v = warp_sum(bq[q]*bq[p]); S[q][p] = ((q < e && p < e) ? H[TRI(32+q, 32+p)] : (q == p ? 1.0 : 0.0)) - v; }
    y2[q] = (q < e ? x[32+q] : 0.0) - warp_sum(bq[q]*b); }
  #pragma unroll
  for (int k = 0; k < E; k++) { S[k][k] = sqrt(fmax(S[k][k], MYO_MINVAL)); double ik = 1.0/S[k][k];
    #pragma unroll
    for (int i = k+1; i < E; i++) S[i][k] *= ik;
    #pragma unroll
    for (int i = k+1; i < E; i++) {
      #pragma unroll
      for (int j = k+1; j <= i; j++) S[i][j] -= S[i][k]*S[j][k]; } }
  #pragma unroll
  for (int k = 0; k < E; k++) { y2[k] /= S[k][k];
    #pragma unroll
    for (int i = k+1; i < E; i++) y2[i] -= S[i][k]*y2[k]; }
  #pragma unroll
  for (int k = E-1; k >= 0; k--) { y2[k] /= S[k][k];
    #pragma unroll
    for (int i = 0; i < k; i++) y2[i] -= S[k][i]*y2[k]; }
  // x1 = L11^-T (y1 - L21' x2)
  #pragma unroll
  for (int q = 0; q < E; q++) b = fma(-bq[q], y2[q], b);
  #pragma unroll
  for (int k = 31; k >= 0; k--) { double xk = __shfl_sync(FULL, b*dinv, k); b = lane == k ? xk : fma(-c[k], xk, b); }
  __syncwarp();
  x[lane] = b;
  #pragma unroll
  for (int q = 0; q < E; q++) if (lane == q && q < e) x[32+q] = y2[q];
  __syncwarp();
}

// x <- H^-1 x for a dense SPD H (packed lower triangle in shared memory); returns whether H survived (the n > 32 fallback factors in place)
__device__ __forceinline__ bool chol_dense(double* H, int n, double* x, int lane, int mode) {
  if (n > 36) { chol_factor_rows(H, n, lane); chol_solve(H, n, x, lane); return false; }
  if (n > 32) { chol_reg32b<4>(H, n, x, lane); return true; }
  if (mode == 0) { chol_smem32(H, n, x, lane); return false; }
  if (n <= 8) chol_reg8(H, n, x, lane); else if (n <= 16) chol_reg16(H, n, x, lane); else if (n <= 24) chol_reg24(H, n, x, lane);
  else chol_reg32(H, n, x, lane);
  return true; }

__device__ __forceinline__ void load_M_dense(const DevModel& m, const Warp& w, double* H, double diag_scale /* h */) {
  int n = m.nv; for (int i = w.lane; i < n*(n+1)/2; i += 32) H[i] = 0; __syncwarp();
  const idx_t* mi = CI(PM_i); const idx_t* mj = CI(PM_j); const double* dofp = CD(PDOF_d);
  for (int e = w.lane; e < m.nM; e += 32) { int i = mi[e], j = mj[e]; double v = w.qM[e]; if (i == j) v += diag_scale*dofp[2*i+1]; H[TRI(i,j)] = v; }
  __syncwarp(); }

// ------------------------------------------------------------------ tree-sparse L'DL (level-scheduled, left-looking) on the qM layout
// Hs (nM, input) -> LD (nM: D on the diagonal slots, unit-L off-diagonals), Dinv (nv)
__device__ void ldl_factor(const DevModel& m, const Warp& w, const double* Hs, double* LD, double* Dinv) {
  const idx_t* fadr = CI(PFE_adr); const idx_t* fe = CI(PFE); const idx_t* tadr = CI(PFT_adr); const idx_t* ft = CI(PFT);
  const idx_t* mi = CI(PM_i); const idx_t* mj = CI(PM_j); const idx_t* madr = CI(dof_Madr);
  for (int lev = m.ndepth-1; lev >= 0; lev--) {
    for (int t = fadr[lev] + w.lane; t < fadr[lev+1]; t += 32) { int e = fe[t]; double v = Hs[e];
      for (int q = tadr[t]; q < tadr[t+1]; q++) v -= LD[ft[3*q]]*LD[ft[3*q+1]]*LD[madr[ft[3*q+2]]];
      LD[e] = v; }
    __syncwarp();
    for (int t = fadr[lev] + w.lane; t < fadr[lev+1]; t += 32) { int e = fe[t], k = mi[e];
      if (k == mj[e]) Dinv[k] = 1.0/LD[e]; else LD[e] /= LD[madr[k]]; }
    __syncwarp(); }
}
// x <- (L'DL)^-1 x
__device__ void ldl_solve(const DevModel& m, const Warp& w, const double* LD, const double* Dinv, double* x) {
  const idx_t* ladr = CI(PLV_adr); const idx_t* lv = CI(PLV); const idx_t* dadr = CI(PDS_adr); const idx_t* ds = CI(PDS);
  const idx_t* madr = CI(dof_Madr); const idx_t* mj = CI(PM_j);
  for (int lev = m.ndepth-1; lev >= 0; lev--) {
    for (int t = ladr[lev] + w.lane; t < ladr[lev+1]; t += 32) { int j = lv[t]; double s = x[j];
      for (int q = dadr[j]; q < dadr[j+1]; q++) s -= LD[ds[2*q+1]]*x[ds[2*q]];
      x[j] = s; }
    __syncwarp(); }
  for (int lev = 0; lev < m.ndepth; lev++) {
    for (int t = ladr[lev] + w.lane; t < ladr[lev+1]; t += 32) { int i = lv[t]; double s = x[i]*Dinv[i];
      for (int e = madr[i]+1; e <= madr[i]+lev; e++) s -= LD[e]*x[mj[e]];
      x[i] = s; }
    __syncwarp(); }
}

// ------------------------------------------------------------------ Newton solver: leaves qacc in s.a
// Called by EVERY warp of the CTA when cta_sync is set (live = this warp holds an env): the warps still iterating are re-aligned by CTA
// barriers at the top of each Newton iteration and before the Cholesky, so that they keep sharing instruction fetches inside the
// phase too (a lone 23x23 register Cholesky costs 29 k cycles out of step with the other warps, 21 k in step); finished and idle
// warps only take part in the barriers.  The total wait is unchanged: the CTA leaves the phase with its slowest env either way.
__device__ void phase_solve(const DevModel& m, Warp& w, double tol, long long* cyc, bool live, bool cta_sync) {
  long long tc = cyc ? clock64() : 0;
  #define LAP(k) if (cyc) { long long t_ = clock64(); cyc[k] += t_ - tc; tc = t_; }
  Solv s = solv_views(m, w); int n = m.nv, nefc = live ? w.nefc : 0; w.niter = 0;
  bool active = live;
  if (live && nefc == 0) {   // unconstrained: qacc = M^-1 qfrc_smooth
    ldl_factor(m, w, w.qM, s.LD, s.Dinv);
    for (int i = w.lane; i < n; i += 32) { s.a[i] = w.fsm[i]; s.Ma[i] = w.fsm[i]; } __syncwarp();
    ldl_solve(m, w, s.LD, s.Dinv, s.a); active = false; }
  if (active) {
    for (int i = w.lane; i < n; i += 32) s.a[i] = w.qws[i]; __syncwarp();
    mul_M(m, w, s.Ma, s.a); rows_apply(m, w, s, s.a, s.jar); __syncwarp();
    for (int r = w.lane; r < nefc; r += 32) s.jar[r] -= s.aref[r]; __syncwarp(); }
  const double scale = 1.0/(m.meaninertia*(n > 1 ? n : 1));
  const idx_t* eq = CI(PEQ); const idx_t* pr = CI(PPAIR); const double* pd = CD(PPAIR_d); const idx_t* path = CI(PPATH);
  for (int iter = 0; iter < 50; iter++) {
    if (cta_sync) { if (!__syncthreads_or(active ? 1 : 0)) break; } else if (!active) break;
    bool dense = false;
    if (active) {
    // gradient
    for (int i = w.lane; i < n; i += 32) s.g[i] = s.Ma[i]-w.fsm[i];
    for (int r = w.lane; r < nefc; r += 32) { double x = s.jar[r]; s.jv[r] = (r < m.neq || x < 0) ? s.D[r]*x : 0.0; }   // jv used as scratch weights
    __syncwarp(); rows_applyT_add(m, w, s, s.jv, s.g); __syncwarp();
    double gn = 0; for (int i = w.lane; i < n; i += 32) gn += s.g[i]*s.g[i]; gn = sqrt(warp_sum(gn));
    LAP(8)
    if (scale*gn < tol) active = false; }
    if (active) {
    if (cyc) cyc[14]++;
    // Hessian: tree-sparse L'DL when no contact row is active (limits/equalities keep M's sparsity), dense Cholesky otherwise
    dense = (m.neq > 0 && !m.eq_tree);
    for (int c = w.lane; c < w.ncon && !dense; c += 32) { int nr = s.cnrow[c], rb = s.crow[c]; for (int r = 0; r < nr; r++) if (s.jar[rb+r] < 0) dense = true; }
    dense = __any_sync(FULL, dense);
    for (int i = w.lane; i < n; i += 32) s.p[i] = -s.g[i];
    __syncwarp();
    if (cyc && dense) cyc[15]++;
    if (!dense) {
      const idx_t* madr = CI(dof_Madr);
      for (int e = w.lane; e < m.nM; e += 32) s.Hs[e] = w.qM[e];
      __syncwarp();
      if (w.lane == 0) for (int e = 0; e < m.neq; e++) { int d1 = eq[PEQ_ISTRIDE*e+1], d2 = eq[PEQ_ISTRIDE*e+3]; double De = s.D[e], j2 = s.eqJ[e]; s.Hs[madr[d1]] += De;
        if (d2 >= 0) { s.Hs[eq[PEQ_ISTRIDE*e+4]] += De*j2; s.Hs[madr[d2]] += De*j2*j2; } }
      __syncwarp();
      for (int pass = 0; pass < 2; pass++) { for (int r = w.lane; r < w.nlimrow; r += 32) { int dsc = s.lrow[r]; if (((dsc >> 16) & 1) == pass && s.jar[m.neq+r] < 0) s.Hs[madr[dsc & 0xffff]] += s.D[m.neq+r]; } __syncwarp(); }
      LAP(9)
      ldl_factor(m, w, s.Hs, s.LD, s.Dinv); ldl_solve(m, w, s.LD, s.Dinv, s.p);
      LAP(10)
    } else {
    load_M_dense(m, w, s.H, 0.0);
    if (w.lane == 0) for (int e = 0; e < m.neq; e++) { int d1 = eq[PEQ_ISTRIDE*e+1], d2 = eq[PEQ_ISTRIDE*e+3]; double De = s.D[e], j2 = s.eqJ[e]; s.H[TRI(d1,d1)] += De;
      if (d2 >= 0) { s.H[d1 > d2 ? TRI(d1,d2) : TRI(d2,d1)] += De*j2; s.H[TRI(d2,d2)] += De*j2*j2; } }
    __syncwarp();
    for (int pass = 0; pass < 2; pass++) { for (int r = w.lane; r < w.nlimrow; r += 32) { int dsc = s.lrow[r]; if (((dsc >> 16) & 1) == pass && s.jar[m.neq+r] < 0) { int d = dsc & 0xffff; s.H[TRI(d,d)] += s.D[m.neq+r]; } } __syncwarp(); }
    for (int c = 0; c < w.ncon; c++) { int nr = s.cnrow[c]; if (!nr) continue; int rb = s.crow[c]; const idx_t* q = pr + PPAIR_ISTRIDE*s.cpair[c]; double W[6] = {0,0,0,0,0,0};  // nn n1 n2 11 12 22
      if (nr == 1) { if (s.jar[rb] < 0) W[0] = s.D[rb]; }
      else { const double* P = pd + q[6]*PPAIR_STRIDE; double mu1 = P[2], mu2 = P[3], Dv = s.D[rb];
        double a0 = s.jar[rb] < 0 ? Dv : 0, a1 = s.jar[rb+1] < 0 ? Dv : 0, a2 = s.jar[rb+2] < 0 ? Dv : 0, a3 = s.jar[rb+3] < 0 ? Dv : 0;
        W[0] = a0+a1+a2+a3; W[1] = mu1*(a0-a1); W[2] = mu2*(a2-a3); W[3] = mu1*mu1*(a0+a1); W[5] = mu2*mu2*(a2+a3); }
      if (W[0] != 0) { const double* J = s.conJ + (size_t)c*3*m.maxpath; int np = q[4], ntri = (np*(np+1)) >> 1;
        // one lane per entry of the lower triangle of J'WJ (the path lists dofs in ascending order, so entry (ei >= ej) lands on H[di >= dj])
        for (int t = w.lane; t < ntri; t += 32) { int ei = __float2int_rd((sqrtf(8.0f*t + 1.0f) - 1.0f)*0.5f); if (((ei+1)*(ei+2) >> 1) <= t) ei++; if (((ei*(ei+1)) >> 1) > t) ei--;
          int ej = t - ((ei*(ei+1)) >> 1);
          const double* a = J + 3*ei; const double* b = J + 3*ej;
          double wa0 = W[0]*a[0]+W[1]*a[1]+W[2]*a[2], wa1 = W[1]*a[0]+W[3]*a[1]+W[4]*a[2], wa2 = W[2]*a[0]+W[4]*a[1]+W[5]*a[2];
          int di = path[q[3]+ei] >> 1, dj = path[q[3]+ej] >> 1; s.H[TRI(di,dj)] += wa0*b[0]+wa1*b[1]+wa2*b[2]; } }
      __syncwarp(); }
    LAP(9)
    } }
    if (cta_sync) __syncthreads();
    if (active) {
    if (dense) { chol_dense(s.H, n, s.p, w.lane, m.chol_mode); LAP(10) }
    // exact line search along p
    mul_M(m, w, s.Mp, s.p); rows_apply(m, w, s, s.p, s.jv); __syncwarp();
    double ga = 0, gb = 0; for (int i = w.lane; i < n; i += 32) { ga += s.p[i]*(s.Ma[i]-w.fsm[i]); gb += s.p[i]*s.Mp[i]; } ga = warp_sum(ga); gb = warp_sum(gb);
    double alpha = 0, lo = 0, hi = -1, d0 = 0;
    for (int it = 0; it < 40; it++) { double dv = 0, hh = 0;
      for (int r = w.lane; r < nefc; r += 32) { double x = s.jar[r]+alpha*s.jv[r]; if (r < m.neq || x < 0) { dv += s.D[r]*x*s.jv[r]; hh += s.D[r]*s.jv[r]*s.jv[r]; } }
      dv = warp_sum(dv) + ga + gb*alpha; hh = warp_sum(hh) + gb;
      if (it == 0) { d0 = fabs(dv); if (dv >= 0) break; } else { if (dv < 0) lo = alpha; else hi = alpha; if (fabs(dv) <= 1e-10*d0) break; }
      double an = alpha - dv/hh;
      if (it > 0 && (an <= lo || (hi > 0 && an >= hi))) an = hi > 0 ? 0.5*(lo+hi) : 2*alpha+1;
      if (an == alpha) break;
      alpha = an; }
    if (alpha == 0) active = false;   // no descent possible: converged to round-off
    else {
    for (int i = w.lane; i < n; i += 32) { s.a[i] += alpha*s.p[i]; s.Ma[i] += alpha*s.Mp[i]; }
    // rows that switch between active and inactive along the step; with none the cost was exactly quadratic along p, the Newton
    // step (alpha = 1) lands on its minimum and the gradient vanishes to round-off: converged without another gradient pass
    bool flip = false;
    for (int r = w.lane; r < nefc; r += 32) { double x0 = s.jar[r], x1 = x0 + alpha*s.jv[r]; s.jar[r] = x1; if (r >= m.neq && (x0 < 0) != (x1 < 0)) flip = true; }
    __syncwarp(); w.niter = iter+1; LAP(11)
    if (!__any_sync(FULL, flip)) active = false; } } }
  #undef LAP
}

// ------------------------------------------------------------------ semi-implicit Euler with implicit joint damping
__device__ void phase_integrate(const DevModel& m, Warp& w, long long* cyc) {
  Solv s = solv_views(m, w); int n = m.nv; double h = m.timestep; long long tc = cyc ? clock64() : 0;
  // (M + h B) qacc' = M qacc  (= qfrc_smooth + qfrc_constraint at the solver optimum)
  for (int i = w.lane; i < n; i += 32) s.g[i] = s.Ma[i];     // M qacc, maintained by the solver
  __syncwarp();
  if (n >= 8 && n <= 36) {     // mid-size systems: the register Cholesky beats the level-scheduled sparse factorisation (long index-chasing chains)
    load_M_dense(m, w, s.H, h);
    if (cyc) { long long t_ = clock64(); cyc[18] += t_ - tc; tc = t_; }
    chol_dense(s.H, n, s.g, w.lane, m.chol_mode);
  } else {
  { const idx_t* mi = CI(PM_i); const idx_t* mj = CI(PM_j); const double* dofp = CD(PDOF_d);
    for (int e = w.lane; e < m.nM; e += 32) { double v = w.qM[e]; if (mi[e] == mj[e]) v += h*dofp[2*mi[e]+1]; s.Hs[e] = v; } __syncwarp(); }
  if (cyc) { long long t_ = clock64(); cyc[18] += t_ - tc; tc = t_; }
  ldl_factor(m, w, s.Hs, s.LD, s.Dinv); ldl_solve(m, w, s.LD, s.Dinv, s.g);
  }
  if (cyc) { long long t_ = clock64(); cyc[19] += t_ - tc; tc = t_; }
  for (int i = w.lane; i < n; i += 32) { w.qvel[i] += h*s.g[i]; w.qws[i] = s.a[i]; }
  __syncwarp();
  const idx_t* jtype = CI(jnt_type); const idx_t* jq = CI(jnt_qposadr); const idx_t* jd = CI(jnt_dofadr);
  for (int j = w.lane; j < m.njnt; j += 32) { int qa = jq[j], da = jd[j];
    if (jtype[j] == 0) { for (int c = 0; c < 3; c++) w.qpos[qa+c] += h*w.qvel[da+c];
      double wv[3] = {w.qvel[da+3], w.qvel[da+4], w.qvel[da+5]}, nn = sqrt(dot3(wv,wv)), ang = h*nn;
      if (nn < MYO_MINVAL) { wv[0]=1; wv[1]=0; wv[2]=0; } else { wv[0]/=nn; wv[1]/=nn; wv[2]/=nn; }
      double sn, cs; sincos(0.5*ang, &sn, &cs); double ql[4] = {cs, wv[0]*sn, wv[1]*sn, wv[2]*sn}, qn[4]; quat_mul(qn, w.qpos+qa+3, ql); quat_norm(qn);
      for (int c = 0; c < 4; c++) w.qpos[qa+3+c] = qn[c]; }
    else w.qpos[qa] += h*w.qvel[da]; }
  __syncwarp();
}
