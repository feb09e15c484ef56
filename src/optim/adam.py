import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ps_pytorch_b200.optim.adam import Adam  # noqa: F401,E402
