"""sieve_b200 -- B200-native tree-likelihood hot path of SIEVE behind the reference's LikelihoodCore seam."""
