"""Test infrastructure only: CPU restatement of the reference's DGDM hot path (dgdm_oracle.py, builder_oracle.py),
the loader for the real reference used in the build container (ref_loader.py) and the golden-vector generator
(make_golden.py). Nothing in the product package imports this."""
