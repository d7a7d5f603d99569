// Test infrastructure: prints the reference's one sprite (make_tile_floor(), sprites.hpp:73-364)
// and its palette (sprites.hpp:60-65) as raw bytes, so the procedurally generated tables of
// the oracle and of the product can be compared with the real ones.  Compiled by
// oracle/build_ref.sh against /root/reference/src (header included from where it lies).
#include <cstdio>

#include "sprites.hpp"

int main() {
    static const Sprite s = make_tile_floor();
    fwrite(&s, sizeof s, 1, stdout);
    fwrite(color_palette, sizeof(Color), 4, stdout);
    return 0;
}
