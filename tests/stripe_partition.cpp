// Host harness around the stripe arithmetic the render kernel and the C ABI share
// (pixel-art-raytracer_b200/csrc/par_device.cuh: owned_tile_rows, stripe_column_segment, stripe_of_segment).
// TEST INFRASTRUCTURE: built (nvcc -x cu, host code only) and run by tests/test_bands_gloo.py.
//
//   stripe_partition W H ranks split      prints one line "rank row0 row1 col0 col1" per owned stripe and
//                                         exits 1 if stripe_of_segment is not the inverse of the deal
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "par_device.cuh"

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]), n = atoi(argv[3]), s = atoi(argv[4]);
    int bad = 0;
    for (int r = 0; r < n; r++) {
        par::ViewDims d{};
        d.W = W;
        d.H = H;
        d.HW = W / par::kBin;
        d.HH = H / par::kBin;
        d.row0 = 0;
        d.row1 = H;
        d.stripe_n = n;
        d.stripe_i = r;
        d.stripe_s = s > 1 ? s : 1;
        d.stripe_rot = std::lcm(d.stripe_n, d.stripe_s);
        int first, count;
        par::owned_tile_rows(d, first, count);
        const int tps = par::tiles_per_stripe(d), seg = par::stripe_segments(d);
        for (int q = 0, v = first; q < count; q++, v += n) {
            const int t = v / seg, c = par::stripe_column_segment(d, v);
            printf("%d %d %d %d %d\n", r, t * par::kBin, (t + 1) * par::kBin, c * tps * par::kBin,
                   seg > 1 ? (c + 1) * tps * par::kBin : W);
            if (par::stripe_of_segment(d, t, c) != v) bad++;
        }
    }
    return bad ? 1 : 0;
}
