"""yelprecommendation_b200 — B200-native (sm_100a) BPR-MF / NGCF training and full-catalog evaluation hot path,
behind the interfaces of twndus/YelpRecommendation (models/mf.py, models/ngcf.py, loss.py, metric.py,
trainers/mf_trainer.py, trainers/ngcf_trainer.py). Compute lives in csrc/*.cu behind include/yelprec_b200.h."""
__version__ = "0.1.0"
