"""Per-folder ``model.py`` shims: put one of these on ``sys.path`` as ``model`` and the
reference's ``train_*.py`` / ``test_*.py`` / ``inference.py`` pick up the B200 generator unchanged."""
