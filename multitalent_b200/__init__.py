"""multitalent_b200 -- B200-native (sm_100a) hot path of MultiTalent behind the reference's module / trainer API."""
__version__ = "0.1.0"
