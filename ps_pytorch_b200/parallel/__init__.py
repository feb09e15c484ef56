from .arena import ParamLayout, FlatBuffer  # noqa: F401
from .transport import Transport, GlooTransport, LocalTransport  # noqa: F401
from .ps import SyncReplicasMaster_NN, GradientAccumulator  # noqa: F401
from .worker import DistributedWorker, ModelBuffer  # noqa: F401
from .evaluator import DistributedEvaluator  # noqa: F401
from .trainer import NN_Trainer  # noqa: F401
from .compute import WorkerCompute  # noqa: F401
