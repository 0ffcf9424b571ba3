"""Import the UNMODIFIED reference (`/root/reference/model.py`) in the build container.

TEST INFRASTRUCTURE.  The reference imports five packages at module top that are not installed
here (model.py:6,9-13; metrics.py:4-7): torchinfo, mlflow, matplotlib(.pyplot),
torchmetrics.functional.image, skimage.metrics.  None of them is on the hot path, so they are
replaced by inert `types.ModuleType` stubs.  Used only by
`oracle/gen_golden.py`; `/root/reference` does not exist on the GPU box, so nothing under
`tests/ -m gpu`, `smoke()` or `bench.py` may call this.
"""
import contextlib
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _noop(*a, **k):
    return None


def install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("torchinfo", summary=lambda *a, **k: "summary-stub")

    @contextlib.contextmanager
    def start_run(*a, **k):
        yield None

    mod("mlflow", log_params=_noop, log_param=_noop, log_metric=_noop, log_metrics=_noop,
        log_artifact=_noop, set_experiment=_noop, start_run=start_run)

    class _AnyAttr(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return _noop

    mpl = _AnyAttr("matplotlib")
    plt = _AnyAttr("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt

    tm = mod("torchmetrics")
    tmf = mod("torchmetrics.functional")
    tmfi = mod("torchmetrics.functional.image", peak_signal_noise_ratio=_noop,
               structural_similarity_index_measure=_noop, spectral_angle_mapper=_noop)
    tm.functional = tmf
    tmf.image = tmfi
    sk = mod("skimage")
    skm = mod("skimage.metrics", peak_signal_noise_ratio=_noop, structural_similarity=_noop)
    sk.metrics = skm


def import_reference_model(root=REFERENCE_ROOT):
    """`root` = /root/reference in the build container, or oracle/_ref (the build-time copy made by oracle/make_ref.py,
    git-ignored, shipped to the GPU box with the built libraries) for bench.py's reference arm."""
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import model  # noqa: the reference's model.py, unmodified
    return model
