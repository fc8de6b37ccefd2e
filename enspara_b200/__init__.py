"""enspara_b200 -- B200-native drop-in for enspara's conformational-clustering hot path.

Importing the package does not touch the GPU; the CUDA library (libenspara_b200.so, built by
``python -m enspara_b200.build``) is loaded on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"
