"""pgx *core* restated from pgx v2's published behaviour (core.py: Env.init / Env.step / Env.observe);
SURVEY.md Appendix B.1.  The env-specific _init/_step/_observe come from the reference files."""
import abc

import jax
import jax.numpy as jnp

from ._src.struct import dataclass

EnvId = str
Array = object


@dataclass
class State(abc.ABC):
    current_player: object
    observation: object
    rewards: object
    terminated: object
    truncated: object
    legal_action_mask: object
    _step_count: object


class Env(abc.ABC):
    def __init__(self):
        pass

    def init(self, key):
        state = self._init(key)
        observation = self.observe(state)
        return state.replace(observation=observation)

    def step(self, state, action, key=None):
        is_illegal = ~state.legal_action_mask[action]
        current_player = state.current_player
        # already-terminal: same state, zero rewards
        state = jax.lax.cond(
            (state.terminated | state.truncated),
            lambda: state.replace(rewards=jnp.zeros_like(state.rewards)),
            lambda: self._step(state.replace(_step_count=state._step_count + 1), action, key),
        )
        state = jax.lax.cond(is_illegal, lambda: self._step_with_illegal_action(state, current_player), lambda: state)
        state = jax.lax.cond(
            state.terminated,
            lambda: state.replace(legal_action_mask=jnp.ones_like(state.legal_action_mask)),
            lambda: state,
        )
        observation = self.observe(state)
        return state.replace(observation=observation)

    def observe(self, state, player_id=None):
        if player_id is None:
            player_id = state.current_player
        return jax.lax.stop_gradient(self._observe(state, player_id))

    @property
    def num_actions(self):
        state = self.init(jax.random.PRNGKey(0))
        return int(state.legal_action_mask.shape[0])

    def _step_with_illegal_action(self, state, loser):
        penalty = -1.0
        reward = jnp.ones_like(state.rewards) * (-1 * penalty) * (self.num_players - 1)
        reward = reward.at[loser].set(penalty)
        return state.replace(rewards=reward, terminated=jnp.bool(True))
