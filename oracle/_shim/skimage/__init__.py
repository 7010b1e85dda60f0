"""Import shim so the reference's ``image_filtering`` module can be imported
where scikit-image is absent (TEST INFRASTRUCTURE ONLY; used by
oracle/make_golden.py).  Not scikit-image."""
