def resource_scatter_add(handle, indices, updates):
    raise NotImplementedError("sparse path is not on the 1-N hot path")
