"""recbole_cdr_b200 -- B200-native hot path for cross-domain recommender training behind the RecBole-CDR model API.

The package mirrors the part of ``recbole_cdr`` that sits on the per-batch path:
``model.cross_domain_recommender.{emcdr,cmf,conet,dtcdr,bitgcf}``, ``trainer.CrossDomainTrainer``,
``sampler.CrossDomainSourceSampler`` and the ``Interaction`` batch dict, all running on ``lib/libxdr.so``
(hand-written sm_100a CUDA behind the C ABI of ``include/xdr.h``).  There is no CPU fallback.
"""
__version__ = '0.1.0'

from . import _lib  # noqa: F401  (fails loudly if libxdr.so is missing)
