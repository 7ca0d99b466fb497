"""cosmicfishpie_b200 -- B200-native observable->Fisher / likelihood hot path of cosmicfishpie."""

__version__ = "0.1.0"
