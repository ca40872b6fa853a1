"""q_learning_with_hjb_b200 — B200-native (sm_100a) implementation of the data-parallel hot path of
HaoxiangYou/Q_Learning_with_HJB: batched closed-loop rollouts of control-affine systems and the vhjb
HJB-residual pass, behind the reference's Dynamics / Controller class interface.  CUDA-only: no CPU fallback."""
__version__ = "0.1.0"
