"""Stub of the un-vendored third-party dependency ``recbole==1.0.1`` (TEST INFRASTRUCTURE ONLY).

RecBole-CDR pins ``recbole==1.0.1`` (reference ``requirements.txt:1``) but the wheel is not in
this image and there is no network.  This package exists so that ``oracle/make_golden.py`` can
import the *unmodified* reference model files from ``/root/reference`` and execute them.

* ``recbole.model.*`` holds REAL restatements of the leaf arithmetic the reference's hot path calls
  (BPRLoss, EmbLoss, RegLoss, MLPLayers, xavier_normal_initialization, AbstractRecommender) --
  restated from the published recbole 1.0.1 semantics (third-party table).
* Everything else is a placeholder that only has to exist because ``recbole_cdr/__init__.py:5``
  imports the whole package (config, data, trainer); none of it is executed by the oracle.

Nothing in the product (``recbole-cdr_b200/``) may import this package.
"""
__version__ = "1.0.1-shim"
