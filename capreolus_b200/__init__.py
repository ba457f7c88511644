"""capreolus_b200 -- B200-native scoring engine behind the Capreolus Reranker API."""
