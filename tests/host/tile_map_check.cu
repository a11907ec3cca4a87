// Host-side check of the accumulator-tile bookkeeping of the moment deposition (xpic_b200/csrc/deposit.cuh): the
// compile-time map "tile element -> (node, stencil slot)" that k_gather_tiles uses is compared with the footprint of a
// particle as the reference deposits it (src/impls/ecsim/particles.cpp:119-171: component c is interpolated with the
// staggered weights along its own axis, base node cell - 1 + octant bit, and with the nodal weights, base node = cell,
// along the other two).  Runs on the CPU; built by xpic_b200/csrc/Makefile, run by tests/test_abi_and_host.py.
#include <cstdio>
#include <set>

#include "../../xpic_b200/csrc/deposit.cuh"

using namespace xb;

static int fails = 0;
#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      if (fails < 20) {             \
        std::printf("FAIL: ");      \
        std::printf(__VA_ARGS__);   \
        std::printf("\n");          \
      }                             \
      ++fails;                      \
    }                               \
  } while (0)

// node offset (relative to the particle's cell) of corner t of component c along axis a for octant bits o[3]
static int footprint(int c, int t, int a, const int* o) { return (a == c ? -1 + o[a] : 0) + ((t >> a) & 1); }

int main()
{
  // 1. tile ids: a bijection onto 0 .. NTILE - 2, the current tile last
  std::set<int> ids;
  int ntiles = 0;
  for (int sl = 0; sl < 9; ++sl)
    for (int v = 0; v < nvar(sl); ++v)
      for (int oz = 0; oz < (dep(sl, 2) ? 2 : 1); ++oz) {
        ids.insert(tile_id(sl, v, oz));
        ++ntiles;
      }
  CHECK(ntiles == NTILE - 1 && (int)ids.size() == ntiles && *ids.begin() == 0 && *ids.rbegin() == NTILE - 2, "tile ids are not a bijection (%d tiles, %zu ids)", ntiles,
        ids.size());
  CHECK(TILE_CUR == NTILE - 1 && NTILE * 64 == STAGE_CELL, "tile count / staging size");

  // 2. every tile element against the particle footprint, for every octant the variant stands for
  long contributions = 0;
  for (int c1 = 0; c1 < 3; ++c1)
    for (int c2 = 0; c2 < 3; ++c2) {
      const int sl = c1 * 3 + c2;
      std::set<int> slots;
      for (int oct = 0; oct < 8; ++oct) {
        const int o[3] = {oct & 1, (oct >> 1) & 1, oct >> 2};
        const int v = vidx(sl, oct & 3) - vbase(sl);  // register variant of this octant's (ox, oy)
        CHECK(v >= 0 && v < nvar(sl), "variant index of slot %d octant %d", sl, oct);
        CHECK(vbit(sl, v, c1, o[2]) == o[c1] && vbit(sl, v, c2, o[2]) == o[c2], "octant bits of slot %d variant %d", sl, v);
        for (int gq = 0; gq < 8; ++gq)
          for (int t2 = 0; t2 < 8; ++t2) {
            const TileTarget tt = tile_target(c1, c2, v, o[2], gq, t2);
            const int n1[3] = {footprint(c1, gq, 0, o), footprint(c1, gq, 1, o), footprint(c1, gq, 2, o)};
            const int n2[3] = {footprint(c2, t2, 0, o), footprint(c2, t2, 1, o), footprint(c2, t2, 2, o)};
            const int d[3] = {n2[0] - n1[0], n2[1] - n1[1], n2[2] - n1[2]};
            CHECK(in_range(c1, c2, d[0], d[1], d[2]), "offset outside the stencil: pair %d %d octant %d row %d column %d", c1, c2, oct, gq, t2);
            // the gather thread of node n reads the cell n - (ox, oy, oz): the row's node is the cell + n1
            CHECK(tt.ox == n1[0] && tt.oy == n1[1] && tt.oz == n1[2], "row node: pair %d %d octant %d row %d: (%d %d %d) against (%d %d %d)", c1, c2, oct, gq, tt.ox,
                  tt.oy, tt.oz, n1[0], n1[1], n1[2]);
            CHECK(tt.slot == coef_slot(c1, c2, d[0], d[1], d[2]), "slot: pair %d %d octant %d row %d column %d", c1, c2, oct, gq, t2);
            CHECK(tt.ox >= -1 && tt.ox <= 1 && tt.oy >= -1 && tt.oy <= 1 && tt.oz >= -1 && tt.oz <= 1, "row node outside the 3 x 3 x 3 cells around the node");
            slots.insert(tt.slot);
          }
      }
      CHECK((int)slots.size() == pair_size(c1, c2) && *slots.begin() == pair_base(c1, c2), "pair %d %d: %zu of %d slots are fed", c1, c2, slots.size(),
            pair_size(c1, c2));
      contributions += 64L * nvar(sl) * (dep(sl, 2) ? 2 : 1);
    }
  CHECK(contributions == 64L * (NTILE - 1), "contributions per cell");

  // 3. the current tile: lane (gq, q = c) holds component c at corner gq for octant bit 0 / 1 of c's own axis
  for (int c = 0; c < 3; ++c)
    for (int ob = 0; ob < 2; ++ob)
      for (int gq = 0; gq < 8; ++gq) {
        int o[3] = {0, 0, 0};
        o[c] = ob;
        const int p = corner_pos(c, gq, ob);
        const int off[3] = {win_lo(c, 0) + pos_i(c, p), win_lo(c, 1) + pos_j(c, p), win_lo(c, 2) + pos_k(c, p)};
        for (int a = 0; a < 3; ++a) CHECK(off[a] == footprint(c, gq, a, o), "current: component %d bit %d corner %d axis %d", c, ob, gq, a);
      }

  if (fails) {
    std::printf("%d checks failed\n", fails);
    return 1;
  }
  std::printf("tile map: %d tiles, %ld contributions per cell, all slots fed: OK\n", NTILE - 1, contributions);
  return 0;
}
