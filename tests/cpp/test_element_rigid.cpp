// The written-out global stiffness of a beam without pins (element.cuh: rigid_global_values / _expand) against the
// literal sequence it shortens (stiff_beam2d + rotate_to_global, feBeam2d.go:22-172, model.go:619-620): every one of
// the 36 entries bit for bit (zeros compared as values: the sign of a zero never reaches the assembled matrix),
// for axis-aligned, diagonal and random directions.  Host build with -ffp-contract=off.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../hd_b200/csrc/element.cuh"

static uint64_t bits(double v) { uint64_t u; std::memcpy(&u, &v, 8); return u; }
static uint64_t rng = 0x9e3779b97f4a7c15ull;
static double uni() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; }

int main()
{
    using namespace hdb;
    const double dirs[][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}, {3, 4}, {-3, 4}, {1, 1}, {-1, -1}, {6, 3}, {1e-9, 1}, {5, -12}};
    long checked = 0;
    for (int it = 0; it < 20000; it++) {
        double dx, dy;
        if (it < 11) { dx = dirs[it][0]; dy = dirs[it][1]; }
        else { dx = 20 * uni() - 10; dy = 20 * uni() - 10; }
        const double x0 = (it % 3 == 0) ? 0.0 : 7 * uni(), y0 = (it % 3 == 0) ? 0.0 : 7 * uni();
        const double x1 = x0 + dx, y1 = y0 + dy;
        const double L = beam_length(x0, y0, x1, y1);
        if (L == 0) continue;
        double c, s;
        beam_cs(x0, y0, x1, y1, L, c, s);
        const BeamCoef cf = beam_coef(12e-4 * (0.5 + 1.5 * uni()), 120e-6 * (0.5 + 1.5 * uni()), 2.0e11, L);
        double k[36], ref[36], d[8], got[36];
        stiff_beam2d(cf, 0, k);
        rotate_to_global(k, c, s, ref);
        rigid_global_values(cf, c, s, d);
        rigid_global_expand(d, got);
        for (int e = 0; e < 36; e++) {
            const bool same = (ref[e] == 0.0 && got[e] == 0.0) || bits(ref[e]) == bits(got[e]);
            if (!same) {
                std::printf("FAIL it %d entry %d: %.17g vs %.17g (c %.17g s %.17g)\n", it, e, ref[e], got[e], c, s);
                return 1;
            }
            // the position table: entry e is +-d[i] for exactly the i whose mask holds bit e
            int owner = -1;
            for (int i = 0; i < 8; i++)
                if ((HD_RIGID_POS(i) >> e) & 1ull) { if (owner >= 0) { std::printf("FAIL mask overlap %d\n", e); return 1; } owner = i; }
            const bool neg = (HD_RIGID_NEG >> e) & 1ull; // the sign table of the compact element store
            if (owner < 0 || got[e] != (neg ? -d[owner] : d[owner])) { std::printf("FAIL mask entry %d\n", e); return 1; }
            checked++;
        }
    }
    std::printf("OK %ld\n", checked);
    return 0;
}
