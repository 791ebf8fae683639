"""Data config of the reference (sqair/configs/seq_mnist_data.py:24-29): two path flags and the `load` of mnist_tools."""
from sqair_b200 import tf_flags as flags
from sqair_b200.mnist_tools import load  # noqa: F401

flags.DEFINE_string('train_path', 'seq_mnist_train.pickle', '')
flags.DEFINE_string('valid_path', 'seq_mnist_validation.pickle', '')
