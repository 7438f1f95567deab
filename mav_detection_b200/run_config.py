"""RunConfig with the reference's API surface (/root/reference/src/run_config.py:13-146): the Mode and
DatasetType enums (1-tuple values, as in the reference), key parsing with the same ValueError texts,
uses_nn_for_detection and the results dict.  Dataset classes (file I/O, ffmpeg, FlowNet2 dockers) are out
of scope; get_dataset() builds one through a factory registry instead of importing them."""
from __future__ import annotations

import json
import logging
import os
from enum import Enum
from typing import Any, Callable, Dict, Iterator, List, Optional, cast


class RunConfig:
    class Mode(Enum):
        APPEARANCE_RGB = 0,
        FLOW_UV = 1,
        FLOW_RADIAL = 2,
        FLOW_FOE_YOLO = 3,
        FLOW_FOE_CLUSTERING = 4,

        def __str__(self) -> str:
            return super().__str__().replace('Mode.', '')

    class DatasetType(Enum):
        MIDGARD = 0,
        SIMULATION = 1,
        EXPERIMENT = 2,
        VIS_DRONE = 3,

        def __str__(self) -> str:
            return super().__str__().replace('DatasetType.', '')

    # DatasetType -> callable(logger, sequence) returning an object with the Dataset accessors the hot loop
    # uses.  The reference hard-wires its four classes here (run_config.py:114-129).
    dataset_factories: Dict['RunConfig.DatasetType', Callable[[logging.Logger, str], Any]] = {}

    @classmethod
    def register_dataset(cls, dataset_type: 'RunConfig.DatasetType', factory: Callable[[logging.Logger, str], Any]) -> None:
        cls.dataset_factories[dataset_type] = factory

    @classmethod
    def get_settings(cls) -> Dict[str, Any]:
        """settings.json from the working directory (run_config.py:33-36); empty when absent."""
        if not os.path.exists('settings.json'):
            return {'train_sequences': [], 'validation_sequences': []}
        with open('settings.json', 'r') as f:
            return cast(Dict[str, Any], json.load(f))

    def __init__(self, logger: logging.Logger, dataset: str, sequence: str, debug: bool, prepare_dataset: bool,
                 validate: bool, headless: bool, data_to_yolo: bool, undistort: bool, mode: str):
        self.logger = logger
        self.dataset = dataset
        self.sequence = sequence
        self.debug = debug
        self.prepare_dataset = prepare_dataset
        self.validate = validate
        self.headless = headless
        self.data_to_yolo = data_to_yolo
        self.undistort = undistort
        self.mode = self.get_mode(mode)
        self.results: dict = dict()
        self.settings = RunConfig.get_settings()

    def get_all_sequences(self) -> List[str]:
        sequences = self.settings['train_sequences']
        for seq in self.settings['validation_sequences']:
            sequences.append(seq)
        return cast(List[str], sequences)

    def uses_nn_for_detection(self) -> bool:
        return self.mode in [RunConfig.Mode.FLOW_UV, RunConfig.Mode.FLOW_RADIAL, RunConfig.Mode.FLOW_FOE_YOLO]

    def get_mode(self, mode_key: str) -> 'RunConfig.Mode':
        options = [mode.name for mode in RunConfig.Mode]
        if mode_key not in options:
            options_str = ', '.join(options)
            raise ValueError(f'Mode {mode_key} is not a valid mode type, has to be one of {options_str}')
        return RunConfig.Mode[mode_key]

    def get_dataset_type(self, dataset_key: str) -> 'RunConfig.DatasetType':
        options = [mode.name for mode in RunConfig.DatasetType]
        dataset_key = dataset_key.upper()
        if dataset_key not in options:
            options_str = ', '.join(options)
            raise ValueError(f'Dataset {dataset_key} is not a valid dataset type, has to be one of {options_str}')
        return RunConfig.DatasetType[dataset_key]

    def get_dataset(self) -> Any:
        data_type = self.get_dataset_type(self.dataset)
        factory = RunConfig.dataset_factories.get(data_type)
        if factory is None:
            raise ValueError(f'Invalid dataset type: {data_type}.')
        dataset = factory(self.logger, self.sequence)
        self.sequence = dataset.sequence
        return dataset

    def __str__(self) -> str:
        return f'{self.dataset}/{self.sequence}/{self.mode}'

    def __iter__(self) -> Iterator[Any]:
        return iter([self.dataset, self.sequence, self.debug, self.prepare_dataset, self.validate, self.headless,
                     self.data_to_yolo, self.undistort, self.mode, *self.results])
