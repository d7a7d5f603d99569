// tile_one_light.cu — the render kernel of tile.cu built a second time, configured for one-light frames
// (6 CTAs per SM, the smallest shared lists; see the top of tile.cu).  Nothing else lives here.
#define PAR_TILE_ONE_LIGHT 1
#include "tile.cu"
