"""Host-side types of the hot path: a minimal mirror of the pieces of ppq.core that the quantization functions and
observers consume (/root/reference/ppq/core/quant.py:123-186, 309-364, 367-896 and core/common.py:10-30).

Names, enum values and property semantics are the reference's so that call sites read the same; nothing of the graph IR,
dominance (union-find) or export logic is mirrored -- that is outside the hot path.
When the real `ppq` package is importable, its own TensorQuantizationConfig objects work with every function here too
(duck typing on the same attribute names).
"""
from enum import Enum
from typing import Optional

import torch

# ppq/core/common.py:10-30
OBSERVER_MIN_SCALE = 1e-8
OBSERVER_MIN_SCALE_MANUL_OVERRIDE = 'OBSERVER_MIN_SCALE_MANUL_OVERRIDE'
OBSERVER_KL_HIST_BINS = 4096
OBSERVER_KL_HIST_BINS_MANUL_OVERRIDE = 'OBSERVER_KL_HIST_BINS_MANUL_OVERRIDE'
OBSERVER_PERCENTILE = 0.9999
OBSERVER_PERCENTILE_MANUL_OVERRIDE = 'OBSERVER_PERCENTILE_MANUL_OVERRIDE'
OBSERVER_MSE_HIST_BINS = 2048
OBSERVER_MSE_COMPUTE_INTERVAL = 8
OBSERVER_FLOATING_MSE_FETCHES = 4096


class RoundingPolicy(Enum):
    """ppq/core/quant.py:123-142 (device ids: ppq/csrc/cuda/common.cuh:17-24; 7 = ROUND_DOWN exists on the device only)."""
    ROUND_HALF_EVEN = 0
    ROUND_HALF_UP = 1
    ROUND_HALF_DOWN = 2
    ROUND_HALF_TOWARDS_ZERO = 3
    ROUND_HALF_FAR_FORM_ZERO = 4
    ROUND_TO_NEAR_INT = 5
    ROUND_UP = 6


class QuantizationProperty(Enum):
    """ppq/core/quant.py:145-186."""
    PER_TENSOR = 0x00000001
    PER_CHANNEL = 0x00000002
    LINEAR = 0x00000004
    FLOATING = 0x00000008
    SYMMETRICAL = 0x00000010
    ASYMMETRICAL = 0x00000020
    POWER_OF_2 = 0x00000040
    DYNAMIC = 0x00000080


class QuantizationPolicy:
    def __init__(self, policy: int):
        self._policy = int(policy)

    def has_property(self, prop: QuantizationProperty) -> bool:
        return (self._policy & prop.value) != 0

    def __eq__(self, o):
        return isinstance(o, QuantizationPolicy) and o._policy == self._policy

    def __hash__(self):
        return hash(self._policy)


class QuantizationStates(Enum):
    """ppq/core/quant.py:309-364."""
    INITIAL = 1
    BAKED = 2
    OVERLAPPED = 3
    ACTIVATED = 4
    PASSIVE = 5
    PASSIVE_INIT = 6
    PASSIVE_BAKED = 7
    FP32 = 8

    @classmethod
    def is_activated(cls, state) -> bool:
        # compared by name so that configs of the real `ppq` package (its own enum class) are understood too
        return getattr(state, 'name', None) in ('ACTIVATED', 'PASSIVE')


def state_is(config, name: str) -> bool:
    """config.state == <name>, for our TensorQuantizationConfig and for the reference's (a different Enum class with the same names)."""
    return getattr(config.state, 'name', None) == name


def set_state(config, name: str) -> None:
    """Assign the member `name` of whatever Enum class the config's state belongs to."""
    config.state = type(config.state)[name]


class TensorQuantizationConfig:
    """The parameter block of every quantization call (ppq/core/quant.py:367-896, ctor :518-599), including the union-find
    `dominated_by` / `master_by` links (:646-712): a dominated config reads scale / offset from the root of its group."""

    def __init__(self, policy: QuantizationPolicy, rounding: RoundingPolicy = RoundingPolicy.ROUND_HALF_EVEN,
                 num_of_bits: int = 8, quant_min=-127, quant_max=128, exponent_bits: int = 0,
                 scale: Optional[torch.Tensor] = None, offset: Optional[torch.Tensor] = None,
                 observer_algorithm: Optional[str] = None, detail: Optional[dict] = None,
                 channel_axis: Optional[int] = None, state: QuantizationStates = QuantizationStates.INITIAL):
        self.policy = policy
        self.rounding = rounding
        self.num_of_bits = num_of_bits
        self.quant_min = quant_min
        self.quant_max = quant_max
        self.exponent_bits = exponent_bits
        self._scale = scale
        self._offset = offset
        self.observer_algorithm = observer_algorithm
        self.detail = {} if detail is None else detail
        self.channel_axis = channel_axis
        self.state = state
        self._dominator = self

    @property
    def mantissa_bits(self) -> int:
        # ppq/core/quant.py:793-800
        return self.num_of_bits - self.exponent_bits - 1

    def is_same_scheme(self, o: 'TensorQuantizationConfig') -> bool:
        """quant.py:632-644."""
        return (self.quant_max == o.quant_max and self.quant_min == o.quant_min and self.policy == o.policy and
                self.num_of_bits == o.num_of_bits and self.exponent_bits == o.exponent_bits and
                self.channel_axis == o.channel_axis and self.rounding == o.rounding)

    # -- union-find over quantization groups (quant.py:646-712)
    @property
    def dominated_by(self) -> 'TensorQuantizationConfig':
        if self._dominator is self: return self
        root = self._dominator.dominated_by
        self._dominator = root                                      # path compression, as upstream
        return root

    @dominated_by.setter
    def dominated_by(self, o: 'TensorQuantizationConfig'):
        assert isinstance(o, TensorQuantizationConfig), 'Can only set this attribute with another tensor config.'
        if o is self: raise ValueError('Error with TQC.dominated_by = o: o must not equal to TQC its self.')
        root, dominator = self.dominated_by, o.dominated_by
        if self is dominator:
            raise ValueError('Can not Assign Dominator like this, Circular reference was detected. Son TQC can not dominate its Father.')
        if dominator is not root:
            root._dominator = dominator
            self._dominator = dominator
            root.state = QuantizationStates.OVERLAPPED
            self.state = QuantizationStates.OVERLAPPED

    @property
    def master_by(self) -> 'TensorQuantizationConfig':
        return self.dominated_by

    @master_by.setter
    def master_by(self, master: 'TensorQuantizationConfig'):
        if not isinstance(master, TensorQuantizationConfig):
            raise TypeError(f'Error with TQC.master_by(o): o must be another Tensor Quantization Config, however {type(master)} was given.')
        if master is self: raise ValueError('Error with TQC.dominated_by = o: o must not equal to TQC its self.')
        self._dominator = master
        self.state = QuantizationStates.PASSIVE if (master.scale is not None and master.offset is not None) else QuantizationStates.PASSIVE_INIT

    def is_revisable(self) -> bool:
        return self.dominated_by is self and self.state in {QuantizationStates.ACTIVATED, QuantizationStates.FP32, QuantizationStates.INITIAL,
                                                           QuantizationStates.PASSIVE, QuantizationStates.PASSIVE_INIT}

    @property
    def scale(self) -> Optional[torch.Tensor]:
        return self._scale if self.dominated_by is self else self.dominated_by.scale

    @scale.setter
    def scale(self, value):
        if not self.is_revisable():
            raise PermissionError('Can not change scale of this tensor quantization configuration now. '
                                  'It has been overlapped or has an inactive state. '
                                  'Due to it is not a active config, any change of this configuration is not allowed.')
        self._scale = value

    @property
    def offset(self) -> Optional[torch.Tensor]:
        return self._offset if self.dominated_by is self else self.dominated_by.offset

    @offset.setter
    def offset(self, value):
        if not self.is_revisable():
            raise PermissionError('Can not change offset of this tensor quantization configuration now. '
                                  'It has been overlapped or has an inactive state. '
                                  'Due to it is not a active config, any change of this configuration is not allowed.')
        self._offset = value


def LinearQuantizationConfig(symmetrical: bool = True, dynamic: bool = False, power_of_2: bool = False,
                             channel_axis: int = None, quant_min: int = -128, quant_max: int = 127, num_of_bits=8,
                             calibration: str = 'minmax',
                             rounding: RoundingPolicy = RoundingPolicy.ROUND_HALF_EVEN) -> TensorQuantizationConfig:
    """ppq/lib/quant.py:106-136."""
    P = QuantizationProperty
    bits = (P.SYMMETRICAL.value if symmetrical else P.ASYMMETRICAL.value) | (P.DYNAMIC.value if dynamic else 0) | \
           (P.POWER_OF_2.value if power_of_2 else 0) | (P.PER_TENSOR.value if channel_axis is None else P.PER_CHANNEL.value) | \
           P.LINEAR.value
    return TensorQuantizationConfig(policy=QuantizationPolicy(bits), rounding=rounding, num_of_bits=num_of_bits,
                                    quant_min=quant_min, quant_max=quant_max, observer_algorithm=calibration,
                                    channel_axis=channel_axis)


def FloatingQuantizationConfig(symmetrical: bool = True, power_of_2: bool = True, channel_axis: int = None,
                               quant_min: float = -448.0, quant_max: float = 448.0, exponent: int = 4, mantissa: int = 3,
                               calibration: str = 'constant',
                               rounding: RoundingPolicy = RoundingPolicy.ROUND_HALF_EVEN) -> TensorQuantizationConfig:
    """ppq/lib/quant.py:139-164 (which drops channel_axis on the floor; kept here because FloatingQuantize_C needs it)."""
    P = QuantizationProperty
    bits = (P.SYMMETRICAL.value if symmetrical else P.ASYMMETRICAL.value) | (P.POWER_OF_2.value if power_of_2 else 0) | \
           (P.PER_TENSOR.value if channel_axis is None else P.PER_CHANNEL.value) | P.FLOATING.value
    return TensorQuantizationConfig(policy=QuantizationPolicy(bits), rounding=rounding, num_of_bits=exponent + mantissa + 1,
                                    exponent_bits=exponent, quant_min=quant_min, quant_max=quant_max,
                                    observer_algorithm=calibration, channel_axis=channel_axis)
