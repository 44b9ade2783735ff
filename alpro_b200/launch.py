"""Run an UNCHANGED reference script on the B200 classes:

    python -m alpro_b200.launch src/tasks/run_video_retrieval.py --config config_release/msrvtt_ret.json ...
    torchrun --nproc-per-node 8 -m alpro_b200.launch src/pretrain/run_pretrain_sparse.py --config ...

(from the reference checkout's root, as run_scripts/*.sh do with horovodrun). Installs the stand-ins of
alpro_b200.shims — horovod.torch over torch.distributed/NCCL, apex.amp pass-through, and
`src.modeling.alpro_models` -> alpro_b200.modeling — then executes the script as __main__.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.exit("usage: python -m alpro_b200.launch <reference script.py> [script args...]")
    script = argv[0]
    from . import shims
    shims.install(alias_models=True, optional_stubs=True)
    here = os.getcwd()
    if here not in sys.path:
        sys.path.insert(0, here)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
