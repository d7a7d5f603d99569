#!/usr/bin/env bash
# Build the REAL reference (Cons-Cat/Pixel-Art-Raytracer) headless, for pinning the oracle.
# TEST INFRASTRUCTURE ONLY.  Sources are compiled where they lie under /root/reference;
# nothing from them is copied into this repo.  Outputs (binaries only) go to oracle/_ref/,
# which is git-ignored but travels to the GPU box.
#
#   tier 0: the unmodified translation unit + oracle/sdl_stub
#   tier 1: the same translation unit streamed through sed into g++'s stdin (no copy is
#           stored) with (a) the three view constants at alternative.cpp:117-119 turned into
#           -D macros, (b) the scene generator (alternative.cpp:519-599) and the light
#           (alternative.cpp:626) pinned to the default 480/320/320 so that only the VIEW
#           grows, (c) a dump hook before draw_line (alternative.cpp:762), (d) a scene hook just
#           before the frame loop (alternative.cpp:628) that, when PAR_REF_SCENE is set, replaces
#           scene and lights by a file's through the reference's own Entities::insert — so the
#           real reference can render arbitrary (random, ragged, multi-light) scenes.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${PAR_REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
src="$ref/src/alternative.cpp"
if [ ! -f "$src" ]; then
    echo "build_ref: $src not present (expected on the GPU box); keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$out"
CXX="$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++ )"
FLAGS="-std=gnu++20 -O3 -w -I$here/sdl_stub"

# tier 0 — unmodified
$CXX $FLAGS "$src" -o "$out/ref_tier0"

# the reference's sprite + palette as bytes on stdout (16000 + 16 bytes)
$CXX $FLAGS -I"$ref/src" "$here/ref_sprite_dump.cpp" -o "$out/ref_sprite_dump"

tier1() { # W H L
    sed -e '117s/= 480;/= PAR_VIEW_W;/' \
        -e '118s/= 320;/= PAR_VIEW_H;/' \
        -e '119s/= 320;/= PAR_VIEW_L; constexpr int scene_width = 480, scene_height = 320, scene_length = 320;/' \
        -e '519,599s/view_width/scene_width/g;519,599s/view_height/scene_height/g;519,599s/view_length/scene_length/g' \
        -e '626s/view_width/scene_width/g;626s/view_height/scene_height/g;626s/view_length/scene_length/g' \
        -e '628s|^|par_stub_load_scene(p_entities, lights);\n|' \
        -e '762s|^|par_stub_dump_pre(p_texture, sizeof(Color) * view_width * view_height, p_pixel_buffer, sizeof(Pixel) * view_width * view_height);\n|' \
        "$src" |
        $CXX $FLAGS -DPAR_VIEW_W=$1 -DPAR_VIEW_H=$2 -DPAR_VIEW_L=$3 -I"$ref/src" -x c++ - -o "$out/ref_tier1_$1x$2x$3"
}
tier1 480 320 320
tier1 1920 1080 1080
tier1 3840 2160 2160
# view length != view height, small views: pins for the oracle's generalisation to runtime W/H/L
tier1 480 320 640
tier1 640 480 200
tier1 200 40 40
tier1 40 1000 120
echo "build_ref: built $(ls "$out" | tr '\n' ' ')"
