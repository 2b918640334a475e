# Mirrors the reference's top-level package `utils`; sub-modules not provided here resolve to a
# reference checkout further down sys.path (see cnhead/_dropin.py).
from cnhead._dropin import extend_package as _extend

__path__ = _extend(__path__, __name__)
