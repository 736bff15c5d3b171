"""Import shim (test infrastructure only) for `omegaconf`: DictConfig is the identity
on plain dicts and OmegaConf.load is yaml.safe_load — all the reference needs
(evaluations/infer_arvc.py:14,53,68,88,99,112)."""
import yaml


def DictConfig(d):
    return d


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return yaml.safe_load(f)
