"""Overlay support: lets this package sit *on top of* a reference checkout.

When UNIDISTILL_REFERENCE_ROOT points at a CVPR2023-UniDistill checkout, every sub-package of this
mirror appends the matching reference directory to its __path__.  Modules that exist here (the hot
path: layers, voxelization, ...) win; everything else (exps/, data/multisensorfusion/, ...) resolves to
the reference's own files, so its experiment scripts run unchanged on the sm_100a kernels.
Without the variable the mirror is self-contained.
"""
import os


def extend(path_list, package_name, namespace=None):
    """Appends the reference's directory of this package to its __path__ and, when the package's `namespace` (its
    globals()) is passed, first runs the reference's own __init__.py in it: names the reference exports from the package
    (e.g. `unidistill.utils.DictAction`) stay importable, and everything the mirror defines afterwards wins."""
    root = os.environ.get("UNIDISTILL_REFERENCE_ROOT")
    if not root:
        return
    sub = os.path.join(root, *package_name.split("."))
    if os.path.isdir(sub) and sub not in path_list:
        path_list.append(sub)
        init = os.path.join(sub, "__init__.py")
        if namespace is not None and os.path.isfile(init):
            try:
                with open(init) as f:
                    exec(compile(f.read(), init, "exec"), namespace)
            except ImportError as exc:       # an optional third-party dependency of a module outside the hot path
                import warnings

                warnings.warn("unidistill overlay: %s: %s" % (init, exc))
