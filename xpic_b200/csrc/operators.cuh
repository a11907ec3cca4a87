// operators.cuh -- the constant part of the implicit Maxwell operator, matrix-free.
// M = 2 I + dt^2/2 curl^- curl^+   (matM, src/impls/ecsim/simulation.cpp:544-552)
#pragma once

namespace xb {

// (curl^- curl^+ f)_c at a node; f(comp, ox, oy, oz) reads component comp at the node + offset.
// (CC f)_c = - d_a^- d_a^+ f_c - d_b^- d_b^+ f_c + d_a^- d_c^+ f_a + d_b^- d_c^+ f_b,  {a, b} = axes != c
// cut_below: the node lies on plane 0 of a box whose z boundary is open.  matM is the PRODUCT of the two curl
// matrices, each with its own out-of-box columns dropped (src/utils/operators.cpp:12-43): the B sites of plane -1
// do not exist, so the backward z difference of curl^- loses its lower term -- which is not what zero ghost
// values of f would give (curl^+ f on plane -1 is not zero).
template <class F>
__device__ __forceinline__ double curlcurl(int c, const double* inv_d, F&& f, bool cut_below = false)
{
  double r = 0.0;
  if (!cut_below) {  // every node of a periodic box, and all but plane 0 of an open one
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a == c) continue;
      int ea[3] = {0, 0, 0}, ec[3] = {0, 0, 0};
      ea[a] = 1;
      ec[c] = 1;
      const double lap = (f(c, ea[0], ea[1], ea[2]) - 2.0 * f(c, 0, 0, 0) + f(c, -ea[0], -ea[1], -ea[2])) * (inv_d[a] * inv_d[a]);
      const double mix = ((f(a, ec[0], ec[1], ec[2]) - f(a, 0, 0, 0)) - (f(a, ec[0] - ea[0], ec[1] - ea[1], ec[2] - ea[2]) - f(a, -ea[0], -ea[1], -ea[2]))) *
                         (inv_d[a] * inv_d[c]);
      r += mix - lap;
    }
    return r;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (a == c) continue;
    int ea[3] = {0, 0, 0}, ec[3] = {0, 0, 0};
    ea[a] = 1;
    ec[c] = 1;
    // (d_a^+ f_c)(0) - (d_a^+ f_c)(-e_a)  and  (d_c^+ f_a)(0) - (d_c^+ f_a)(-e_a): along z the second terms sit on plane -1
    const double up_l = f(c, ea[0], ea[1], ea[2]) - f(c, 0, 0, 0);
    const double dn_l = a == 2 ? 0.0 : f(c, 0, 0, 0) - f(c, -ea[0], -ea[1], -ea[2]);
    const double up_m = f(a, ec[0], ec[1], ec[2]) - f(a, 0, 0, 0);
    const double dn_m = a == 2 ? 0.0 : f(a, ec[0] - ea[0], ec[1] - ea[1], ec[2] - ea[2]) - f(a, -ea[0], -ea[1], -ea[2]);
    r += (up_m - dn_m) * (inv_d[a] * inv_d[c]) - (up_l - dn_l) * (inv_d[a] * inv_d[a]);
  }
  return r;
}

}  // namespace xb
