"""Build-time generator: linear-parametric Lindblad generators -> unrolled CUDA matvecs."""
