"""Import alias: `autostyle-tts_b200/` is not a valid Python identifier, so this module loads the
package by path and re-exports it.  `import autostyle_tts_b200 as avs; avs.MilvusClient(...)`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("autostyle-tts_b200")
sys.modules[__name__] = _pkg
