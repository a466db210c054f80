"""Negative samplers of the hot path: device-side uniform draw with rejection against the user's
interaction CSR (csrc/neg_sample.cu), behind the reference's sampler class names and call signatures."""
from .crossdomain_sampler import CrossDomainSourceSampler, TargetDomainSampler, build_used_csr  # noqa: F401
