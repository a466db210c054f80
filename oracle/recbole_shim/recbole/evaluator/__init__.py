metric_types = {}
smaller_metrics = []
