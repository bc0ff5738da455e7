// Host-side check of the building blocks in csrc/edt_fused.cuh: the phases of the fused EDT kernels are executed
// thread after thread (a __syncthreads between phases == finishing the loop) and the result is compared with a
// brute-force squared EDT.  usage: edt_fused_host_test Nx Ny Nz NB W density seed [wide]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../3d-astar-path-planners_b200/csrc/edt_fused.cuh"
using namespace b200::edtf;
static const int kInf = 1 << 29;

static int bits_of(unsigned v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

template <int NB, int W, class TrY, class TrX, bool PACKED>
static int run(int Nx, int Ny, int Nz, double density, unsigned seed, int pillars) {
  const int wz = (Nz + 31) / 32;
  std::mt19937 rng(seed);
  std::vector<uint32_t> occ((size_t)Nx * Ny * wz, 0);
  std::uniform_real_distribution<double> U(0, 1);
  if (pillars) {
    for (int p = 0; p < pillars; ++p) {
      int cx = rng() % Nx, cy = rng() % Ny, w = 1 + rng() % 6, h = 1 + rng() % Nz;
      for (int x = cx; x < cx + w && x < Nx; ++x) for (int y = cy; y < cy + w && y < Ny; ++y) for (int z = 0; z < h; ++z)
        occ[((size_t)x * Ny + y) * wz + (z >> 5)] |= 1u << (z & 31);
    }
  } else {
    for (int x = 0; x < Nx; ++x) for (int y = 0; y < Ny; ++y) for (int z = 0; z < Nz; ++z)
      if (U(rng) < density) occ[((size_t)x * Ny + y) * wz + (z >> 5)] |= 1u << (z & 31);
  }
  const size_t nv = (size_t)Nx * Ny * Nz;
  // ---- reference: separable brute force
  std::vector<int> ref(nv);
  {
    std::vector<int> a(nv), b(nv);
    for (int x = 0; x < Nx; ++x) for (int y = 0; y < Ny; ++y) for (int z = 0; z < Nz; ++z) {
      int best = kInf;
      for (int s = 0; s < Nz; ++s) if (occ[((size_t)x * Ny + y) * wz + (s >> 5)] >> (s & 31) & 1u) { int d = (z - s) * (z - s); if (d < best) best = d; }
      a[((size_t)x * Ny + y) * Nz + z] = best;
    }
    for (int x = 0; x < Nx; ++x) for (int z = 0; z < Nz; ++z) for (int y = 0; y < Ny; ++y) {
      long long best = kInf;
      for (int s = 0; s < Ny; ++s) { long long v = (long long)(y - s) * (y - s) + a[((size_t)x * Ny + s) * Nz + z]; if (v < best) best = v; }
      b[((size_t)x * Ny + y) * Nz + z] = (int)best;
    }
    for (int y = 0; y < Ny; ++y) for (int z = 0; z < Nz; ++z) for (int x = 0; x < Nx; ++x) {
      long long best = kInf;
      for (int s = 0; s < Nx; ++s) { long long v = (long long)(x - s) * (x - s) + b[((size_t)s * Ny + y) * Nz + z]; if (v < best) best = v; }
      ref[((size_t)x * Ny + y) * Nz + z] = (int)best;
    }
  }
  // ---- emulated kernels
  using EY = typename TrY::E;
  using EX = typename TrX::E;
  const int ztiles = (Nz + W - 1) / W;
  const int DYB = bits_of(Ny - 1);
  const int SBy = bits_of((Ny + NB - 1) / NB - 1), SBx = bits_of((Nx + NB - 1) / NB - 1);
  std::vector<uint16_t> tmp16(PACKED ? nv : 0);
  std::vector<int> tmp32(PACKED ? 0 : nv);
  std::vector<uint8_t> plane_empty(Nx, 0);
  std::vector<int> out(nv, -1);
  {  // pass y
    const int BL = (Ny + NB - 1) / NB;
    std::vector<EY> st((size_t)NB * BL * W);
    std::vector<int> below(Ny), above(Ny);
    std::vector<uint32_t> rg(NB * W);
    std::vector<uint16_t> off(NB * W);
    std::vector<int> total(W);
    for (int x = 0; x < Nx; ++x) for (int zt = 0; zt < ztiles; ++zt) {
      const int z0 = zt * W;
      const int wi = z0 >> 5;
      for (int y = 0; y < Ny; ++y) column_outside(&occ[((size_t)x * Ny + y) * wz], wz, Nz, wi, 0, 0, &below[y], &above[y]);
      for (int band = 0; band < NB; ++band) for (int lane = 0; lane < W; ++lane) {
        HullBuilder<TrY> hb;
        hb.init(&st[(size_t)band * BL * W + lane], W, SBy, band * BL);
        if (z0 + lane < Nz)
          for (int u = band * BL; u < Ny && u < (band + 1) * BL; ++u) {
            const int z = z0 + lane, d = site_dz(occ[((size_t)x * Ny + u) * wz + wi], z & 31, z, below[u], above[u]);
            if (d < kDzNone) hb.push(u, d);
          }
        rg[band * W + lane] = hb.range();
      }
      for (int gs = 2; gs <= NB; gs <<= 1)
        for (int band = 0; band < NB; band += gs) for (int lane = 0; lane < W; ++lane) {
          LineHull<TrY> L{&st[lane], &rg[lane], W, SBy, BL};
          zipper<TrY>(L, band, band + gs / 2, band + gs);
        }
      for (int lane = 0; lane < W; ++lane) total[lane] = hull_offsets(&rg[lane], &off[lane], W, NB);
      for (int band = 0; band < NB; ++band) for (int lane = 0; lane < W; ++lane) {
        const int z = z0 + lane;
        if (z >= Nz) continue;
        const int T = total[lane];
        const int u0 = band * BL, u1 = std::min(Ny, (band + 1) * BL);
        if (T == 0) {
          if (PACKED) plane_empty[x] = 1;
          else for (int u = u0; u < u1; ++u) tmp32[((size_t)x * Ny + u) * Nz + z] = kInf;
          continue;
        }
        LineHull<TrY> L{&st[lane], &rg[lane], W, SBy, BL};
        int u = u0;
        hull_fill<TrY>(L, &off[lane], NB, T, u0, u1, (int)off[band * W + lane], [&](uint32_t p) { return p << DYB; }, [&](int d, uint32_t ps, int32_t v) {
          const size_t a = ((size_t)x * Ny + u) * Nz + z;
          if (PACKED) tmp16[a] = (uint16_t)((d < 0 ? -d : d) | ps);
          else tmp32[a] = (int)v;
          ++u;
        });
      }
    }
  }
  {  // pass x
    const int BL = (Nx + NB - 1) / NB;
    std::vector<EX> st((size_t)NB * BL * W);
    std::vector<uint32_t> rg(NB * W);
    std::vector<uint16_t> off(NB * W);
    std::vector<int> total(W);
    for (int y = 0; y < Ny; ++y) for (int zt = 0; zt < ztiles; ++zt) {
      const int z0 = zt * W;
      for (int band = 0; band < NB; ++band) for (int lane = 0; lane < W; ++lane) {
        HullBuilder<TrX> hb;
        hb.init(&st[(size_t)band * BL * W + lane], W, SBx, band * BL);
        const int z = z0 + lane;
        if (z < Nz)
          for (int u = band * BL; u < Nx && u < (band + 1) * BL; ++u) {
            const size_t a = ((size_t)u * Ny + y) * Nz + z;
            if (PACKED) {
              if (plane_empty[u]) continue;
              const int c = tmp16[a], dy = c & ((1 << DYB) - 1), dz = c >> DYB;
              hb.push(u, (uint32_t)(dy * dy + dz * dz));
            } else if (tmp32[a] < kInf) hb.push(u, (uint32_t)tmp32[a]);
          }
        rg[band * W + lane] = hb.range();
      }
      for (int gs = 2; gs <= NB; gs <<= 1)
        for (int band = 0; band < NB; band += gs) for (int lane = 0; lane < W; ++lane) {
          LineHull<TrX> L{&st[lane], &rg[lane], W, SBx, BL};
          zipper<TrX>(L, band, band + gs / 2, band + gs);
        }
      for (int lane = 0; lane < W; ++lane) total[lane] = hull_offsets(&rg[lane], &off[lane], W, NB);
      for (int band = 0; band < NB; ++band) for (int lane = 0; lane < W; ++lane) {
        const int z = z0 + lane;
        if (z >= Nz) continue;
        const int T = total[lane];
        const int u0 = band * BL, u1 = std::min(Nx, (band + 1) * BL);
        if (T == 0) { for (int u = u0; u < u1; ++u) out[((size_t)u * Ny + y) * Nz + z] = kInf; continue; }
        LineHull<TrX> L{&st[lane], &rg[lane], W, SBx, BL};
        int u = u0;
        hull_fill<TrX>(L, &off[lane], NB, T, u0, u1, (int)off[band * W + lane], [](uint32_t) { return 0; }, [&](int, int, int32_t v) { out[((size_t)u * Ny + y) * Nz + z] = (int)v; ++u; });
      }
    }
  }
  size_t bad = 0;
  for (size_t i = 0; i < nv; ++i) if (out[i] != ref[i]) { if (bad < 5) printf("  mismatch at %zu: got %d want %d\n", i, out[i], ref[i]); ++bad; }
  printf("%dx%dx%d NB=%d W=%d packed=%d seed=%u: %zu mismatches\n", Nx, Ny, Nz, NB, W, (int)PACKED, seed, bad);
  return bad != 0;
}

int main(int argc, char** argv) {
  int rc = 0;
  using SY = PassY<uint16_t, int32_t>; using SX = PassX<uint32_t, int32_t>;
  using GY = PassY<uint32_t, long long>; using GX = PassX<uint64_t, long long>;
  using LX = PassX<uint32_t, long long>; using GY2 = PassY<uint16_t, long long>;
  unsigned seed = argc > 1 ? atoi(argv[1]) : 1;
  for (unsigned s = seed; s < seed + 3; ++s) {
    rc |= run<8, 32, SY, SX, true>(40, 70, 50, 0.002, s, 0);
    rc |= run<8, 32, SY, SX, true>(33, 64, 64, 0.05, s, 0);
    rc |= run<8, 32, SY, SX, true>(48, 96, 40, 0.0, s, 9);
    rc |= run<8, 32, SY, SX, true>(20, 20, 33, 0.0, s, 0);   // empty map
    rc |= run<8, 32, SY, SX, true>(64, 5, 7, 0.01, s, 0);    // fewer sites than bands
    rc |= run<8, 32, GY, GX, false>(40, 70, 50, 0.002, s, 0);
    rc |= run<16, 16, GY, GX, false>(50, 90, 40, 0.001, s, 0);
    rc |= run<32, 8, GY, GX, false>(70, 130, 20, 0.0005, s, 0);
    rc |= run<32, 8, GY, GX, false>(30, 60, 24, 0.0, s, 5);
    rc |= run<32, 8, GY, LX, false>(70, 200, 24, 0.0003, s, 0);
    rc |= run<16, 16, GY, LX, false>(100, 50, 36, 0.0, s, 4);
    rc |= run<32, 8, GY2, LX, false>(40, 300, 20, 0.0004, s, 0);
  }
  printf(rc ? "FAILED\n" : "ALL OK\n");
  return rc;
}
