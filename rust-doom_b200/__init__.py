"""b2d -- B200-native software renderer for the Doom-WAD visibility-and-raster hot path.

Host-side mirror of the reference's renderer-facing surface (wad::Archive / WadSystem /
engine::Renderer) over the C-ABI shared library `libb2d.so` (include/b2d.h).  See DESIGN.md.
"""
__version__ = "0.1.0"
