dataset_arguments = []
