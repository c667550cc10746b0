from . import reference
from .functional import set_backend, get_backend, lstm_cell_step, lstm_layer_sequence, head_xent
from .loss import compute_loss, compute_accuracy
