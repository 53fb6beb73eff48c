"""`import mmx` -> the package in ./multimodal-embedding_b200 (its directory name is not a Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("multimodal-embedding_b200")
