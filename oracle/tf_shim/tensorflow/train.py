"""`tf.train` stand-ins for ctr/tfrecord_io.py:66-74 — NOT TensorFlow, test infrastructure only (see io.py)."""


class BytesList:
    def __init__(self, value):
        self.value = list(value)


class Int64List:
    def __init__(self, value):
        self.value = list(value)


class Feature:
    def __init__(self, bytes_list=None, int64_list=None):
        self.value = (bytes_list if bytes_list is not None else int64_list).value


class Features:
    def __init__(self, feature):
        self.feature = dict(feature)


class Example:
    def __init__(self, features):
        self.features = features

    def SerializeToString(self):
        return {k: v.value[0] for k, v in self.features.feature.items()}
