"""Import alias: the package directory name required by the repo layout
(`listening-to-sound-of-silence-for-speech-denoising_b200/`) is not a valid Python identifier, so
`import sos_b200` loads that directory as the package `sos_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "listening-to-sound-of-silence-for-speech-denoising_b200")
_spec = importlib.util.spec_from_file_location("sos_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sos_b200"] = _mod
_spec.loader.exec_module(_mod)
