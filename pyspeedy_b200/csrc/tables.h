// speedy-b200: host table generator interface (tables.cu)
#pragma once
#include "spdy.cuh"
namespace spdy {
void build_tables(ConstTables &C, GlobTables &G);
extern const double H_REARTH, H_OMEGA, H_GRAV, H_P0, H_CP, H_AKAP, H_RGAS, H_GAMMA, H_HSCALE, H_HSHUM, H_DELT;
}  // namespace spdy
