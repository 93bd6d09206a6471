"""nerffaceediting_b200 — B200-native tri-plane volume-rendering hot path of NeRFFaceEditing.

Host side of the drop-in: Python classes mirroring the reference's renderer / ray-marcher /
ray-sampler / decoder interfaces over a C-ABI CUDA library (include/nfe_b200.h).  Importing the
package does not load the library; the first kernel call does, and fails loudly if it is missing.
"""
__version__ = "0.1.0"
