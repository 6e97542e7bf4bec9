"""hifihr_b200 — B200-native (sm_100a) implementation of HiFiHR's hand-mesh render hot path.

Public surface mirrors the reference's names for this path:
  ManoLayer, MyMANOLayer                      (utils/my_mano.py)
  Meshes                                      (pytorch3d.structures, the subset used)
  RasterizationSettings, MeshRasterizer, Fragments, MeshRenderer, HardPhongShader,
  SoftPhongShader, SoftSilhouetteShader, Materials, DirectionalLights, PerspectiveCameras,
  BlendParams, TexturesUV, TexturesUVPCA      (pytorch3d.renderer, the subset used; PCA = NIMBLE texture model)
  LossFunction                                (losses.py, render-dependent terms)
  texture_metrics                             (train_hrnet.py:149-161 PSNR / SSIM / L1 / L2)
  HandRenderModel, FusedHandStep, FusedNimbleStep   (models_res_nimble.py:133-223 without the CNNs)

All compute runs in hand-written CUDA kernels behind the C-ABI of include/hifihr_b200.h
(libhifihr_b200.so, loaded with ctypes).  There is no CPU / PyTorch fallback.
"""
from ._lib import ENTRY_POINTS, LIB_PATH, HfrError  # noqa: F401
from .losses import LossFunction, texture_metrics, trans_proj_j2d  # noqa: F401
from .mano import ManoLayer, MyMANOLayer, xyz_from_vertice  # noqa: F401
from .model import FusedHandStep, FusedNimbleStep, HandRenderModel, get_ndc_fx_fy_cx_cy  # noqa: F401
from .renderer import (BlendParams, DirectionalLights, Fragments, HardPhongShader, Materials,  # noqa: F401
                       MeshRasterizer, MeshRenderer, PerspectiveCameras, PointLights, RasterizationSettings,
                       SoftPhongShader, SoftSilhouetteShader, TexturesUV, TexturesUVPCA, rasterize_meshes)
from .structures import Meshes  # noqa: F401

__version__ = "0.1.0"
