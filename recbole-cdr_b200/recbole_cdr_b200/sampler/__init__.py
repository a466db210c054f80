from .crossdomain_sampler import CrossDomainSourceSampler, TargetDomainSampler, build_used_csr  # noqa: F401
