"""``summarizer`` — the reference's package name, served by summarizer_b200.

``from summarizer.models.vasnet import VASNet``, ``from summarizer.utils.eval import generate_summary``,
``from summarizer.main import train`` ... resolve to the SAME module objects as their ``summarizer_b200.*``
counterparts (a meta-path alias, no second copy of the code), so scripts written against
sylvainma/Summarizer run unchanged with this repository's root on ``sys.path``."""
import importlib
import importlib.abc
import importlib.util
import sys

import summarizer_b200 as _impl

_PREFIX, _REAL = "summarizer.", "summarizer_b200."


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if not name.startswith(_PREFIX):
            return None
        try:
            importlib.import_module(_REAL + name[len(_PREFIX):])
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(name, self)

    def create_module(self, spec):
        return sys.modules[_REAL + spec.name[len(_PREFIX):]]

    def exec_module(self, module):
        pass


if not any(type(f).__name__ == "_AliasFinder" for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
__version__ = _impl.__version__
