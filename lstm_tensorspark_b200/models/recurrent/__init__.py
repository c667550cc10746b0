from .lstm import LSTMLayer, create_variable, truncated_normal_
from .rnn import RNN
