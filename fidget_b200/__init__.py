"""fidget_b200 -- B200-native (sm_100a CUDA) backend for Fidget's tape
evaluation hot path: interval evaluation + tape simplification over tiles and
bulk f32 / gradient evaluation of the surviving voxels.

Layout
  host.py    host-side tape front end (Context, .vm loader, SSA, register
             allocation, bytecode) -- mirrors fidget-core / fidget-bytecode
  shape.py   CudaShape / evaluators / pixel.render / voxel.render -- mirrors
             the reference's Shape + fidget-raster API on top of the C ABI
  effects.py fidget-raster's post-processing effects (denoise, SSAO, shading, RGBA conversions)
  csrc/      CUDA kernels + the C ABI (include/fidget_cuda.h)
"""
from .host import Context, TapeData, Bytecode, OPCODES  # noqa: F401
from .shape import (  # noqa: F401
    CudaContext, CudaShape, CudaError, RenderConfig2D, RenderConfig3D, GEOMETRY_PIXEL,
    render2d, render3d, octree_sample, mesh, schedule_check, OCTREE_LEAF, pixel_inside, screen_to_world_2d, screen_to_world_3d, pixel_mat, voxel_mat,
)
from . import effects  # noqa: F401,E402
