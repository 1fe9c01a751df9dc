import tensorflow as _tf


def variable(initial_value, name=None, trainable=True):
    """Non-trainable scalar state (beta powers). A stored value in state.init_values resumes a previous step."""
    if name in _tf.state.variables:
        return _tf.state.variables[name]
    val = _tf.state.init_values.get(name, initial_value)
    v = _tf.Variable(name, val, trainable=False)
    _tf.state.variables[name] = v
    return v
