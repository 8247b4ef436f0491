// Host emulation of k_normalise_wm's tile algorithm (branch next/weight-map) against the mirror's
// den_from_weight_map: same loops, threads emulated sequentially per phase.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
constexpr int WM_TZ = 4, WM_TY = 8, WM_TX = 32;
constexpr int WM_EZ = WM_TZ + 3, WM_EY = WM_TY + 3, WM_EX = WM_TX + 3;
static void den_ref(const std::vector<int64_t> &G, int D, int H, int W, const float kf[4], std::vector<double> &den) {
    const int64_t V = (int64_t)D * H * W;
    std::vector<double> a(V), b(V);
    const double k[4] = {kf[0], kf[1], kf[2], kf[3]};
    for (int64_t zy = 0; zy < (int64_t)D * H; ++zy)
        for (int x = 0; x < W; ++x) { double acc = 0; for (int d = 0; d < 4 && d <= x; ++d) acc = std::fma(k[d], (double)G[zy * W + x - d], acc); a[zy * W + x] = acc; }
    for (int z = 0; z < D; ++z) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        double acc = 0; for (int d = 0; d < 4 && d <= y; ++d) acc = std::fma(k[d], a[((int64_t)z * H + y - d) * W + x], acc); b[((int64_t)z * H + y) * W + x] = acc; }
    den.assign(V, 0.0);
    for (int z = 0; z < D; ++z) for (int64_t yx = 0; yx < (int64_t)H * W; ++yx) {
        double acc = 0; for (int d = 0; d < 4 && d <= z; ++d) acc = std::fma(k[d], b[(int64_t)(z - d) * H * W + yx], acc); den[(int64_t)z * H * W + yx] = acc; }
}
static void den_tiles(const std::vector<int64_t> &G, int D, int H, int W, int z0, int z1, const float kf[4], std::vector<double> &den) {
    static double sa[WM_EZ][WM_EY][WM_EX], sb[WM_EZ][WM_EY][WM_EX];
    const double k[4] = {kf[0], kf[1], kf[2], kf[3]};
    const int tx = (W + WM_TX - 1) / WM_TX, ty = (H + WM_TY - 1) / WM_TY, tz = (z1 - z0 + WM_TZ - 1) / WM_TZ;
    const long long tiles = (long long)tz * ty * tx;
    for (long long t = 0; t < tiles; ++t) {
        const int ix = (int)(t % tx), iy = (int)((t / tx) % ty), iz = (int)((t / ((long long)tx * ty)) % tz);
        const int X0 = ix * WM_TX, Y0 = iy * WM_TY, Z0 = z0 + iz * WM_TZ;
        for (int i = 0; i < WM_EZ * WM_EY * WM_EX; ++i) {
            const int lx = i % WM_EX, ly = (i / WM_EX) % WM_EY, lz = i / (WM_EX * WM_EY);
            const int gz = Z0 - 3 + lz, gy = Y0 - 3 + ly, gx = X0 - 3 + lx;
            const bool in = (unsigned)gz < (unsigned)D && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
            sa[lz][ly][lx] = in ? (double)G[((long long)gz * H + gy) * W + gx] : 0.0;
        }
        for (int i = 0; i < WM_EZ * WM_EY * WM_TX; ++i) {
            const int lx = 3 + i % WM_TX, ly = (i / WM_TX) % WM_EY, lz = i / (WM_TX * WM_EY);
            double acc = 0.0; for (int d = 0; d < 4; ++d) acc = std::fma(k[d], sa[lz][ly][lx - d], acc); sb[lz][ly][lx] = acc;
        }
        for (int i = 0; i < WM_EZ * WM_TY * WM_TX; ++i) {
            const int lx = 3 + i % WM_TX, ly = 3 + (i / WM_TX) % WM_TY, lz = i / (WM_TX * WM_TY);
            double acc = 0.0; for (int d = 0; d < 4; ++d) acc = std::fma(k[d], sb[lz][ly - d][lx], acc); sa[lz][ly][lx] = acc;
        }
        for (int i = 0; i < WM_TZ * WM_TY * WM_TX; ++i) {
            const int lx = 3 + i % WM_TX, ly = 3 + (i / WM_TX) % WM_TY, lz = 3 + i / (WM_TX * WM_TY);
            const int gz = Z0 + lz - 3, gy = Y0 + ly - 3, gx = X0 + lx - 3;
            if (gz < z1 && gy < H && gx < W) {
                double dn = 0.0; for (int d = 0; d < 4; ++d) dn = std::fma(k[d], sa[lz - d][ly][lx], dn);
                den[((long long)gz * H + gy) * W + gx] = dn;
            }
        }
    }
}
int main() {
    const float kf[4] = {0.43869004f, 0.89384985f, 0.89384985f, 0.43869004f};
    int bad = 0;
    for (auto shp : std::vector<std::vector<int>>{{9, 10, 11}, {21, 30, 17}, {4, 4, 4}, {40, 33, 70}, {13, 8, 32}}) {
        const int D = shp[0], H = shp[1], W = shp[2];
        std::vector<int64_t> G((size_t)D * H * W);
        srand(D * 131 + W);
        for (auto &g : G) g = (rand() % 5 == 0) ? (int64_t)(rand() % (96 << 20)) : 0;
        std::vector<double> a, b((size_t)D * H * W, -1.0);
        den_ref(G, D, H, W, kf, a);
        den_tiles(G, D, H, W, 0, D, kf, b);
        long long diff = 0; for (size_t i = 0; i < a.size(); ++i) diff += (a[i] != b[i]);
        // a plane range, as the chunked stage-2 path issues it
        std::vector<double> c((size_t)D * H * W, -1.0);
        const int z0 = D / 3, z1 = std::max(z0 + 1, 2 * D / 3);
        den_tiles(G, D, H, W, z0, z1, kf, c);
        long long diff2 = 0, untouched = 0;
        for (int z = 0; z < D; ++z) for (int i = 0; i < H * W; ++i) { const size_t j = (size_t)z * H * W + i; if (z >= z0 && z < z1) diff2 += (a[j] != c[j]); else untouched += (c[j] != -1.0); }
        printf("%dx%dx%d: whole mismatches %lld, range mismatches %lld, written outside range %lld\n", D, H, W, diff, diff2, untouched);
        bad += (diff || diff2 || untouched);
    }
    return bad ? 1 : 0;
}
