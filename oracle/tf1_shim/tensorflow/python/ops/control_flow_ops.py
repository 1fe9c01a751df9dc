def group(*ops, **k):
    return list(ops)
