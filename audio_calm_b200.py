"""Import shim: registers the hyphenated directory ``audio-calm_b200/`` as the package ``audio_calm_b200``.

``import audio_calm_b200`` executes this file, which loads ``audio-calm_b200/__init__.py`` under the same module
name (with that directory as the package search path) and replaces itself in ``sys.modules``.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "audio-calm_b200")
_spec = _ilu.spec_from_file_location("audio_calm_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["audio_calm_b200"] = _mod
_spec.loader.exec_module(_mod)
