"""oracle/fixtures.py — the small configurations the golden vectors are generated on.  TEST INFRASTRUCTURE.

Shared by oracle/make_golden.py (runs the real reference), tests/ (oracle + CUDA parity) and smoke().
Channel counts are multiples of 64 so the same fixtures run through the tcgen05 kernels."""
import copy

MINI_ENC = (
    ('causal-conv3d', {'in_channels': 3, 'out_channels': 64, 'kernel_size': 3}),
    ('video-residual', {'in_channels': 64}),
    ('spacetime_downsample', {'in_channels': 64, 'out_channels': 64, 'kernel_size': 3, 'time_factor': 1, 'space_factor': 2}),
    ('video-residual', {'in_channels': 64, 'out_channels': 128}),
    ('spacetime_downsample', {'in_channels': 128, 'out_channels': 128, 'kernel_size': 3, 'time_factor': 2, 'space_factor': 2}),
    ('video-residual', {'n_rep': 2, 'in_channels': 128}),
    ('group_norm', {'num_groups': 8, 'num_channels': 128}),
    ('silu', {}),
    ('causal-conv3d', {'in_channels': 128, 'out_channels': 6, 'kernel_size': 1}),
)

MINI_DEC = (
    ('causal-conv3d', {'in_channels': 6, 'out_channels': 128, 'kernel_size': 3}),
    ('video-residual', {'in_channels': 128}),
    ('adaptive_group_norm', {'dim_cond': 6, 'num_groups': 8, 'num_channels': 128, 'has_ext': True}),
    ('depth2spacetime_upsample', {'in_channels': 128, 'kernel_size': 3, 'time_factor': 2, 'space_factor': 2}),
    ('adaptive_group_norm', {'dim_cond': 6, 'num_groups': 8, 'num_channels': 128, 'has_ext': True}),
    ('video-residual', {'in_channels': 128, 'out_channels': 64}),
    ('depth2spacetime_upsample', {'in_channels': 64, 'kernel_size': 3, 'time_factor': 1, 'space_factor': 2}),
    ('video-residual', {'in_channels': 64}),
    ('group_norm', {'num_groups': 8, 'num_channels': 64}),
    ('silu', {}),
    ('causal-conv3d', {'in_channels': 64, 'out_channels': 3, 'kernel_size': 3}),
)
MINI_D_CODEBOOK = 6
MINI_VIDEO_SHAPE = (2, 3, 8, 32, 32)

# LatentAction, pinned HEAD-valid form of genie/__init__.py:10-54, shrunk
# d_head = 64 like the shipped blueprints (the tcgen05 attention kernel is specialised for it)
MINI_ACT_ENC = (
    ('space-time_attn', {'n_rep': 1, 'n_head': 2, 'd_head': 64, 'transpose': True}),
    ('spacetime_downsample', {'in_channels': 128, 'kernel_size': 3, 'time_factor': 1, 'space_factor': 2}),
    ('space-time_attn', {'n_rep': 1, 'n_head': 2, 'd_head': 64, 'transpose': True}),
)
MINI_ACT_DEC = (
    ('space-time_attn', {'n_rep': 1, 'n_head': 2, 'd_head': 64, 'transpose': True, 'has_ext': True,
                         'time_attn_kw': {'key_dim': 4}}),
    ('depth2spacetime_upsample', {'in_channels': 128, 'kernel_size': 3, 'time_factor': 1, 'space_factor': 2}),
    ('space-time_attn', {'n_rep': 1, 'n_head': 2, 'd_head': 64, 'transpose': True, 'has_ext': True,
                         'time_attn_kw': {'key_dim': 4}}),
)
MINI_ACT_D_CODEBOOK = 4
MINI_ACT_EMBD = 128
MINI_ACT_VIDEO_SHAPE = (2, 3, 4, 16, 16)

MINI_DYN_DESC = (('space-time_attn', {'n_rep': 2, 'n_head': 2, 'd_head': 64, 'transpose': False}),)
MINI_DYN = dict(tok_vocab=64, act_vocab=16, embed_dim=128)
MINI_DYN_TOKENS_SHAPE = (2, 4, 8, 8)

# SpaceTimeAttention block fixtures: (transpose, cond_dim, input shape)
ST_BLOCK_CASES = (
    (True, None, (2, 128, 2, 16, 16)),      # S = 256: two key tiles
    (True, 4, (2, 128, 4, 8, 8)),           # S = 64: masked partial tile; temporal cond K/V
    (False, None, (2, 4, 8, 8, 128)),       # channels-last logical layout (DynamicsModel)
    (False, 4, (2, 4, 8, 8, 128)),
)


def bp(desc):
    """Blueprints must be deep-copied before every parse: the reference's parse_blueprint pops keys
    from the caller's dicts (genie/module/__init__.py:82-88)."""
    return copy.deepcopy(desc)
