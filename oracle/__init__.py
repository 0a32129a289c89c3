"""CPU oracle: a plain torch/numpy fp32 restatement of the reference TF-1.x graph.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The shipped package
(``adaptive-multispeaker-separation_b200/``) never imports this module and fails
loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference (Totoketchup/Adaptive-MultiSpeaker-Separation @
8e7e869) is Python-2 + TensorFlow 1.4/1.5 graph code with no tests, no golden
vectors and no fixtures, and neither Python 2 nor
TensorFlow can be installed in this image, so the oracle cannot be checked
against reference outputs.  All arithmetic lives in the un-vendored dependency
``tensorflow_gpu==1.4.0`` / ``tensorflow==1.5.0rc1`` (requirements.txt:6,11);
each function below restates the published semantics of the TF op the reference
calls and cites the reference call site (file:line under /root/reference).  To
make up for the missing pin every TF-op restatement is cross-checked in
``tests/test_oracle.py`` against an independent formulation (torch.nn.LSTM,
torch.nn.functional.conv1d / conv_transpose1d, scipy.signal, numpy FFT,
finite differences).
"""
from . import tf_ops, models, kmeans, amsgrad  # noqa: F401
