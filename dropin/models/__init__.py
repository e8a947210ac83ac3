"""Shadow of the reference's ``models`` package: put ``dropin/`` first on PYTHONPATH."""
