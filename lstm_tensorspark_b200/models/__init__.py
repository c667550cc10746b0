from .recurrent import LSTMLayer, RNN, create_variable
from .classifier import SequenceClassifier, DenseHead
