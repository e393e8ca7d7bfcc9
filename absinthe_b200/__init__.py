"""absinthe_b200 -- a B200-native execution backend for Absinthe's fused/tiled stencil sequences.

The package holds only what the hot path needs: the analyzer that turns a
variant description into a kernel spec (``analysis``), the stencil catalogue
(``stencils``/``programs``), the CUDA emitter (``emitter``/``generator``) and
the ctypes host for the C-ABI launcher (``launcher``).  There is no CPU
fallback: executing a variant without the CUDA extension raises.
"""

__version__ = "0.1.0"
