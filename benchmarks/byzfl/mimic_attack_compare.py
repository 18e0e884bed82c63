"""Mimic attack: this library vs the external ByzFL library on the same synthetic inputs (counterpart of the
reference's benchmarks/byzfl/mimic_attack_compare.py).  Thin front-end of benchmarks/byzfl/compare.py with the operator
fixed; its flags apply (--num-grads --grad-dim --f --repeat --timeout).  ByzFL is not installable in
the offline build image: the script then reports ``byzfl: unavailable`` next to this library's time.

    python benchmarks/byzfl/mimic_attack_compare.py --num-grads 64 --grad-dim 65536
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchmarks.byzfl.compare import main  # noqa: E402

if __name__ == "__main__":
    if "--op" not in sys.argv:
        sys.argv += ["--op", "mimic"]
    main()
