"""Command-line wrapper of lpbox.train_policy.train_lp_policy (the reference recipe, LP.trainer:254-299, on GPU-produced iterates).
Run:  python tools/train_policy.py [n_instances] [epochs] [out.pt]   (LPBOX_TRAIN_POS_WEIGHT / LPBOX_TRAIN_NEG_WEIGHT: class weights)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
from lpbox.train_policy import train_lp_policy  # noqa: E402

if __name__ == "__main__":
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "accelerated-lpbox-admm_b200", "lpbox", "weights", "lp_mha_policy.pt")
    train_lp_policy(n_inst, epochs, out, pos_weight=float(os.environ.get("LPBOX_TRAIN_POS_WEIGHT", "1")),
                    neg_weight=float(os.environ.get("LPBOX_TRAIN_NEG_WEIGHT", "1")), log=lambda s: print(s, flush=True))
