"""Minimal mpi4py stand-in for running the UNMODIFIED reference without MPI (mpi4py / mpirun are not in this
image — SURVEY §6).  ``MPI.COMM_WORLD`` maps the reference's blocking pickled p2p (``comm.send`` / ``comm.recv``),
``Barrier``, ``Get_rank/Get_size`` and ``Abort`` onto ``torch.distributed`` with the gloo backend (state_dicts
are CPU tensors in the reference — it moves every model to the CPU before sending)."""
from . import MPI  # noqa: F401
