"""Only the two output containers that context.py / selfplay.py construct.  The search itself is NOT here."""
import chex

EpistemicRecurrentFn = object


@chex.dataclass
class EpistemicRecurrentFnOutput:
    reward: object
    reward_epistemic_variance: object
    discount: object
    prior_logits: object
    value: object
    value_epistemic_variance: object


@chex.dataclass
class EpistemicRootFnOutput:
    prior_logits: object
    value: object
    value_epistemic_variance: object
    embedding: object
    beta: object
