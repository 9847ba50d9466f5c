"""Import shim: the product package lives in the directory ``bns-gcn_b200/`` (the
name the build contract fixes); a hyphen is not a legal Python identifier, so
``import bns_gcn_b200`` resolves to this file, which loads that directory as the
package ``bns_gcn_b200`` and replaces itself in ``sys.modules``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bns-gcn_b200")
_spec = importlib.util.spec_from_file_location(
    "bns_gcn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bns_gcn_b200"] = _mod
_spec.loader.exec_module(_mod)
