"""Minimal stand-in for the 12 fairseq symbols FitHuBERT's hot path imports.

TEST INFRASTRUCTURE ONLY. fairseq (pinned by the reference at commit
1b61bbad327d2bf32502b3b9a770b57714cc43dc, reference Dockerfile:10) is not
installed in this image; this package restates the published semantics of the
symbols imported at reference modules/model.py:6-8 and modules/module.py:9-22
so that the *unmodified* reference files can be executed on CPU to pin the
oracle (see oracle/gen_golden.py).  Nothing in fithubert_b200/ imports it.
"""
from . import utils  # noqa: F401
