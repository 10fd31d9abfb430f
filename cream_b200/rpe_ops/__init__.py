"""Drop-in for the reference package `rpe_ops` (iRPE/DeiT-with-iRPE/rpe_ops): put this
directory's parent on sys.path (or `pip install -e`) so that irpe.py's
`from rpe_ops.rpe_index import RPEIndexFunction` (irpe.py:8-15) resolves here."""
