"""B200-native candidate-batched noise-search step (drop-in for rvignav/diffusion-tts' EDM path).

Import as `diffusion_tts_b200` (the root-level shim module maps that name onto this
directory, whose on-disk name carries a hyphen).  Sub-modules:
  build     nvcc build of csrc/ -> libb200ns.so (sm_100a)
  _lib/ops  ctypes binding of the C ABI (include/b200_noise_search.h) and torch-tensor wrappers
  unet      U-Net engine: weight packing + kernel plan for DhariwalUNet / SongUNet
  scorers   Scorer classes with the reference's call protocol
  edm.main  SamplingMethod / SamplingParams / generate_image_grid with the reference's signature
"""
__version__ = '0.1.0'
