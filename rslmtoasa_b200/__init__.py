"""rslmtoasa_b200 -- B200-native engine for the recursion hot path of RS-LMTO-ASA (rslmtoasa/rslmtoasa).

Contents: `csrc/` (CUDA kernels + C ABI -> librsrec.so), `recursion.py` (host-side mirror of the reference's
`type recursion`), `green.py` (mirrors of the `green` / `dos` / `conductivity` consumers), `synthetic.py` (input generators in the reference's data conventions), `build.py`.
"""
from .recursion import Recursion, Control, Energy  # noqa: F401
from ._lib import RsrecError  # noqa: F401
from .green import Green, Dos, Conductivity  # noqa: F401
