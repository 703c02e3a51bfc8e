// speedy-b200: unity build of the CUDA library (one translation unit, so the __constant__ tables are shared
// without relocatable device code).  Build: see pyspeedy_b200/csrc/Makefile.
#include "transforms.cu"
#include "fused_mma3.cu"
#include "fused_mma4.cu"
#include "fused_mma2.cu"
#include "dynamics.cu"
#include "physics.cu"
#include "surface.cu"
#include "sppt.cu"
#include "tables.cu"
#include "engine.cu"
