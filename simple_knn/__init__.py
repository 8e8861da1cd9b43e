"""Drop-in `simple_knn` package: `from simple_knn._C import distCUDA2` (reference scene/gaussian_model.py:20)."""
