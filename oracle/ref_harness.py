"""Import the UNMODIFIED reference (test / bench infrastructure only).

Two locations: /root/reference in the build container, or the byte-identical staged copy
``oracle/_ref/reference`` that ``oracle/build_ref.py`` makes (git-ignored, travels to the GPU box,
where /root/reference does not exist).  Used by ``oracle/gen_golden.py`` to produce the committed
fixtures under ``tests/golden/``, by the tests that pin ``oracle/ynet_oracle.py`` against the live
reference, and by ``bench.py --impl reference`` (the reference's own evaluate() on the host cores).

Recipe follows SURVEY.md Appendix A:
  * ``loralib`` -> ``oracle/loralib_restatement.py`` (package not installed);
  * empty ``seaborn`` / ``matplotlib`` stubs (imported at module top by
    /root/reference/utils/data_utils.py:8-11, never called on the path).
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'reference')


def _find_ref():
    for p in (os.environ.get('REF_PATH'), '/root/reference', _STAGED):
        if p and os.path.isdir(os.path.join(p, 'models')):
            return p
    return os.environ.get('REF_PATH', '/root/reference')


REF_PATH = _find_ref()

_loaded = {}


def available():
    return os.path.isdir(os.path.join(REF_PATH, 'models'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Return a namespace with the reference's modules (imported once).

    The reference uses top-level package names ``models`` and ``utils``; they are
    imported with REF_PATH temporarily at the front of sys.path and then kept in
    sys.modules under their own names (the product package uses the distinct
    name ``motion_style_transfer_b200`` so there is no clash).
    """
    if _loaded:
        return _loaded['ns']
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_PATH}')
    from oracle import loralib_restatement
    loralib_restatement.install_as_loralib()
    mpl = _stub('matplotlib', rcParams={})
    _stub('matplotlib.pyplot')
    mpl.pyplot = sys.modules['matplotlib.pyplot']
    _stub('seaborn')
    sys.path.insert(0, REF_PATH)
    try:
        ns = types.SimpleNamespace()
        ns.image_utils = importlib.import_module('utils.image_utils')
        ns.softargmax = importlib.import_module('utils.softargmax')
        ns.kmeans = importlib.import_module('utils.kmeans')
        ns.evaluate = importlib.import_module('utils.evaluate')
        ns.dataloader = importlib.import_module('utils.dataloader')
        ns.ynet = importlib.import_module('models.ynet')
        try:
            ns.train_epoch = importlib.import_module('utils.train_epoch')
            ns.trainer = importlib.import_module('models.trainer')
        except Exception as e:  # pragma: no cover - depends on optional deps
            ns.train_epoch = None
            ns.trainer = None
            ns.trainer_error = repr(e)
    finally:
        sys.path.remove(REF_PATH)
    _loaded['ns'] = ns
    return ns


def load_scripts_host():
    """``load()`` plus the host side of the reference's scripts: ``utils.data_utils``, ``utils.util``,
    ``utils.parser``, ``utils.extract_log``.  Kept out of ``load()``: the reference arm of bench.py does not need them."""
    ns = load()
    if getattr(ns, 'data_utils', None) is None:
        sys.path.insert(0, REF_PATH)
        try:
            for name in ('data_utils', 'util', 'parser', 'extract_log'):
                setattr(ns, name, importlib.import_module('utils.' + name))
        finally:
            sys.path.remove(REF_PATH)
    return ns
