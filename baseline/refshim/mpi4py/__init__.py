"""Minimal single-host stand-in for the ``mpi4py`` package, used ONLY to run the unmodified reference
(hwang595/ps_pytorch) for ``bench.py --impl reference`` on a box that has neither an MPI library nor network access.

It implements exactly the subset of the mpi4py API the reference calls (see ``MPI.py``); the transport is one Unix-domain
stream socket per ordered (sender, receiver) pair with a reader thread per connection, so sends never block on the peer's
receive calls (the eager protocol an MPI library uses for these message sizes).  It is not part of the product."""
