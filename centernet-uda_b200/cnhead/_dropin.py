"""Glue that lets this tree shadow the reference checkout on ``sys.path``.

Run the unmodified reference with ``PYTHONPATH=<repo>/centernet-uda_b200:<reference>``: the packages
``losses``, ``backends`` and ``utils`` found first are ours; ``extend_package`` makes sub-modules we do
not provide (``backends.dla``, ``utils.helper`` ...) resolve to the reference, and ``reexport`` pulls the
names we do not define (``utils.image.gaussian_radius`` ...) out of the reference module we shadow.
"""
import importlib.util
import os
import sys
from pkgutil import extend_path


def extend_package(path, name):
    return extend_path(path, name)


def reexport(module_name: str, module_file: str, namespace: dict) -> None:
    rel = os.path.join(*module_name.split(".")) + ".py"
    here = os.path.realpath(module_file)
    for entry in sys.path:
        cand = os.path.join(entry or ".", rel)
        if not os.path.isfile(cand) or os.path.realpath(cand) == here:
            continue
        spec = importlib.util.spec_from_file_location("_cnh_shadowed_" + module_name.replace(".", "_"), cand)
        mod = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception:          # the shadowed module needs something that is not installed
            return
        for k, v in vars(mod).items():
            if not k.startswith("__") and k not in namespace:
                namespace[k] = v
        return
