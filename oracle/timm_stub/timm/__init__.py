"""Test-only stand-in for the (absent, unpinned) `timm` dependency of the reference.

Only used by oracle/make_golden.py, in the build container, to import the UNMODIFIED reference
modules from /root/reference.  Never imported by the product package.
"""
