"""Import shim (reference volleyball.py:2-3 imports skimage but the path never uses it)."""
