"""RunConfig with the reference's API surface (/root/reference/src/run_config.py:13-146) for the hot path: the Mode and
DatasetType enums (1-tuple values, printed by bare name, as in the reference), key parsing with the reference's
ValueError texts, uses_nn_for_detection and the results dict that Processor.run_detection fills.  The reference's
dataset classes (file I/O, ffmpeg, FlowNet2 dockers) are out of scope: get_dataset() asks a factory registry for an
object with the accessors the hot loop uses instead of importing them."""
from __future__ import annotations

import json
import logging
import os
from enum import Enum
from typing import Any, Callable, Dict, Iterator, List, Type, TypeVar

E = TypeVar('E', bound=Enum)


class _BareNameEnum(Enum):
    """str(member) is the bare member name (the reference strips the class prefix, run_config.py:21-22, 30-31)."""

    def __str__(self) -> str:
        return self.name


def _parse_key(enum_cls: Type[E], key: str, message: str) -> E:
    names = [member.name for member in enum_cls]
    if key not in names:
        raise ValueError(message % (key, ', '.join(names)))
    return enum_cls[key]


class RunConfig:
    # values are 1-tuples because the reference's members carry trailing commas (run_config.py:15-19, 25-28)
    Mode = _BareNameEnum('Mode', [('APPEARANCE_RGB', (0,)), ('FLOW_UV', (1,)), ('FLOW_RADIAL', (2,)),
                                  ('FLOW_FOE_YOLO', (3,)), ('FLOW_FOE_CLUSTERING', (4,))])
    DatasetType = _BareNameEnum('DatasetType', [('MIDGARD', (0,)), ('SIMULATION', (1,)), ('EXPERIMENT', (2,)),
                                                ('VIS_DRONE', (3,))])

    # modes whose detections come from a neural network on rendered flow images (run_config.py:70-75)
    NN_MODES = ('FLOW_UV', 'FLOW_RADIAL', 'FLOW_FOE_YOLO')

    # DatasetType -> callable(logger, sequence) -> dataset object (the reference hard-wires four classes, :114-129)
    dataset_factories: Dict[Any, Callable[[logging.Logger, str], Any]] = {}

    @classmethod
    def register_dataset(cls, dataset_type: Any, factory: Callable[[logging.Logger, str], Any]) -> None:
        cls.dataset_factories[dataset_type] = factory

    @staticmethod
    def get_settings() -> Dict[str, Any]:
        """settings.json of the working directory (run_config.py:33-36); no sequences when the file is absent."""
        if os.path.exists('settings.json'):
            with open('settings.json', 'r') as handle:
                return dict(json.load(handle))
        return {'train_sequences': [], 'validation_sequences': []}

    def __init__(self, logger: logging.Logger, dataset: str, sequence: str, debug: bool, prepare_dataset: bool,
                 validate: bool, headless: bool, data_to_yolo: bool, undistort: bool, mode: str):
        self.logger, self.dataset, self.sequence = logger, dataset, sequence
        self.debug, self.prepare_dataset, self.validate = debug, prepare_dataset, validate
        self.headless, self.data_to_yolo, self.undistort = headless, data_to_yolo, undistort
        self.mode = self.get_mode(mode)
        self.results: Dict[int, Any] = {}          # frame index -> FrameResult (processor.py:382-383)
        self.settings = RunConfig.get_settings()

    # -- key parsing (same messages as run_config.py:87-91, 106-110) ---------------------------------------------
    def get_mode(self, mode_key: str) -> Any:
        return _parse_key(RunConfig.Mode, mode_key, 'Mode %s is not a valid mode type, has to be one of %s')

    def get_dataset_type(self, dataset_key: str) -> Any:
        return _parse_key(RunConfig.DatasetType, dataset_key.upper(),
                          'Dataset %s is not a valid dataset type, has to be one of %s')

    # -- queries -------------------------------------------------------------------------------------------------
    def uses_nn_for_detection(self) -> bool:
        return self.mode.name in RunConfig.NN_MODES

    def get_all_sequences(self) -> List[str]:
        return list(self.settings['train_sequences']) + list(self.settings['validation_sequences'])

    def get_dataset(self) -> Any:
        kind = self.get_dataset_type(self.dataset)
        if kind not in RunConfig.dataset_factories:
            raise ValueError('Invalid dataset type: %s.' % kind)
        dataset = RunConfig.dataset_factories[kind](self.logger, self.sequence)
        self.sequence = dataset.sequence
        return dataset

    def __str__(self) -> str:
        return '/'.join((self.dataset, self.sequence, str(self.mode)))

    def __iter__(self) -> Iterator[Any]:
        yield from (self.dataset, self.sequence, self.debug, self.prepare_dataset, self.validate, self.headless,
                    self.data_to_yolo, self.undistort, self.mode)
        yield from self.results
