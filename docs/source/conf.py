# Sphinx configuration (MyST markdown sources, autodoc API reference).
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

project = "byzpy_b200"
author = "byzpy_b200 developers"
release = "0.1.0"
extensions = ["sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode", "myst_parser"]
source_suffix = {".rst": "restructuredtext", ".md": "markdown"}
master_doc = "index"
exclude_patterns = ["_build"]
html_theme = "alabaster"
autodoc_mock_imports = ["byzpy_b200._C"]
autodoc_default_options = {"members": True, "undoc-members": False, "show-inheritance": True}
html_static_path = ["_static"]
html_css_files = ["custom.css"]
