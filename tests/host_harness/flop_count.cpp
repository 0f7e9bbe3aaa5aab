// TEST / MEASUREMENT ONLY: exact FP64 operation count of the product's cost-functor evaluation.
// model.cuh (the __host__ __device__ functions the kernels run) is compiled here with `double` replaced by a
// counting scalar, so every +, -, *, /, sqrt, sin the analytic residual + Jacobian evaluation executes is
// counted -- the number bench.py's FP64 roofline uses (SURVEY.md 8(d4) asked for an instrumented count
// instead of an estimate).  Never loaded by the product package.
//   g++ -O1 -std=c++17 tests/host_harness/flop_count.cpp -o flop_count && ./flop_count   -> one JSON line
#include <math.h>
#include <stdio.h>

struct Count { long add = 0, mul = 0, div = 0, sqrt_ = 0, trig = 0; } g_cnt;

struct cd {
  double v;
  cd() : v(0) {}
  cd(double x) : v(x) {}
  cd(int x) : v(x) {}
};
static inline cd operator+(cd a, cd b) { g_cnt.add++; return cd(a.v + b.v); }
static inline cd operator-(cd a, cd b) { g_cnt.add++; return cd(a.v - b.v); }
static inline cd operator*(cd a, cd b) { g_cnt.mul++; return cd(a.v * b.v); }
static inline cd operator/(cd a, cd b) { g_cnt.div++; return cd(a.v / b.v); }
static inline cd operator-(cd a) { return cd(-a.v); }            // sign flip: free (folds into the consumer)
static inline bool operator<(cd a, cd b) { return a.v < b.v; }
static inline bool operator<=(cd a, cd b) { return a.v <= b.v; }
static inline bool operator>(cd a, cd b) { return a.v > b.v; }
static inline bool operator>=(cd a, cd b) { return a.v >= b.v; }
static inline cd& operator+=(cd& a, cd b) { a = a + b; return a; }
static inline cd& operator-=(cd& a, cd b) { a = a - b; return a; }
static inline cd& operator*=(cd& a, cd b) { a = a * b; return a; }
static inline cd sqrt(cd a) { g_cnt.sqrt_++; return cd(::sqrt(a.v)); }
static inline cd sin(cd a) { g_cnt.trig++; return cd(::sin(a.v)); }
static inline cd cos(cd a) { g_cnt.trig++; return cd(::cos(a.v)); }
static inline cd log(cd a) { g_cnt.trig++; return cd(::log(a.v)); }

#define double cd
#include "../../robot_camera_calibration_b200/csrc/model.cuh"
#undef double
using namespace rcc;

template <bool RIG>
static void one_block(Count& expand, Count& geom, Count& corners) {
  cd view6[6] = {0.1, -0.2, 0.05, 0.3, -0.1, -2.0}, marker6[6] = {0.2, 0.1, -0.3, 0.1, 0.2, 0.3};
  cd ext6[6] = {0.01, 0.2, -0.02, 0.05, 0.0, 0.01}, sh9[9] = {600, 610, 320, 240, 0.1, -0.05, 1e-3, -2e-3, 0.01};
  cd vx[POSEX], mx[POSEX], xx[POSEX];
  g_cnt = Count();
  expand_pose(view6, vx);
  expand_marker_pose(marker6, mx);
  if (RIG) expand_pose(ext6, xx);
  expand = g_cnt;
  g_cnt = Count();
  BlockGeom<RIG> g;
  block_geometry<RIG>(vx, mx, RIG ? xx : nullptr, g);
  geom = g_cnt;
  g_cnt = Count();
  for (int k = 0; k < 4; ++k) {
    cd ox, oy;
    corner_xy(k, cd(0.05), ox, oy);
    CornerRows<RIG> c;
    eval_corner<RIG, true>(g, sh9, ox, oy, cd(300.0), cd(200.0), c);
  }
  corners = g_cnt;
}

static long flops(const Count& c) { return c.add + c.mul + c.div + c.sqrt_ + c.trig; }
static void print(const char* name, const Count& c, const char* tail) {
  printf("\"%s\": {\"add\": %ld, \"mul\": %ld, \"div\": %ld, \"sqrt\": %ld, \"trig\": %ld, \"flop\": %ld}%s", name, c.add,
         c.mul, c.div, c.sqrt_, c.trig, flops(c), tail);
}

int main() {
  Count e, g, c, er, gr, cr;
  one_block<false>(e, g, c);
  one_block<true>(er, gr, cr);
  printf("{");
  print("single_expand_2_poses", e, ", ");
  print("single_block_geometry", g, ", ");
  print("single_4_corners", c, ", ");
  print("rig_expand_3_poses", er, ", ");
  print("rig_block_geometry", gr, ", ");
  print("rig_4_corners", cr, ", ");
  // the J^T J / J^T r products are a count, not a measurement: 2 flop x 8 residual rows per distinct entry.
  // single: row = [own 6 | other 6 | shared 9 | r]: own x own 21, own x other 36, own x [shared r] 60,
  // other x other 21, other x [shared r] 60, [shared r] x [shared r] 55 (incl. r.r = the cost)
  const long e_entries = 21 + 36 + 60 + 55, all_entries = e_entries + 21 + 60;
  printf("\"single_products_e_pass\": %ld, \"single_products_all\": %ld, ", 2 * 8 * e_entries, 2 * 8 * all_entries);
  printf("\"single_eval_per_block\": %ld, \"rig_eval_per_block\": %ld}\n", flops(g) + flops(c), flops(gr) + flops(cr));
  return 0;
}
