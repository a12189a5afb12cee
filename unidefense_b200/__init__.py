"""unidefense_b200 -- B200-native (sm_100a) implementation of UniDefense's dual-space
reconstruction hot path behind the reference's model/loss API."""
__version__ = "0.1.0"
