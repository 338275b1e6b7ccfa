"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Minimal stand-in for the third-party ``yolox==0.3.0`` package (pinned at
/root/reference/README.md:67, not vendored in the reference, not installable
here: no network).  Only the nine symbols that ``/root/reference/exps/model/*``
import are restated; it exists so that
``oracle/make_golden.py`` can import the UNMODIFIED reference model files in
this container and generate the fixtures under tests/golden/.
"""
