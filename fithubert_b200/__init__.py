"""fithubert_b200: B200-native (sm_100a) FitHuBERT distillation hot path.

Public surface mirrors the reference: CustomStudentModelConfig / CustomStudentModel (modules/model.py),
TeacherWrapper (utils/utils.py), W2V2Distil (train.py), UpstreamExpert (fithubert/expert.py)."""
from .config import CustomStudentModelConfig, parse_layer_spec  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need CUDA
    if name in ("CustomStudentModel", "TeacherModel", "TeacherWrapper", "freeze_model"):
        from . import model
        return getattr(model, name)
    if name in ("W2V2Distil", "load_model_and_config"):
        from . import distill
        return getattr(distill, name)
    if name in ("UpstreamExpert", "fithubert"):
        from . import expert
        return getattr(expert, name)
    if name in ("load_fairseq_teacher", "load_student_state_dict", "load_checkpoint_to_cpu", "student_checkpoint"):
        from . import checkpoint
        return getattr(checkpoint, name)
    if name in ("LibriDataset", "SyntheticBuckets", "BucketLoader", "shard_indices", "load_audio"):
        from . import data
        return getattr(data, name)
    if name in ("fit", "total_training_steps", "TopKCheckpoints", "EarlyStopping"):
        from . import trainer
        return getattr(trainer, name)
    if name in ("FusedAdamW", "GradAllReduce", "warmup_linear"):
        from . import optim
        return getattr(optim, name)
    raise AttributeError(name)
