from .comm import Communicator, make_communicator
from .launch import launch, RankFailure
