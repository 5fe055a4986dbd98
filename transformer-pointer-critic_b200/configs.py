"""configs/configs.py:5-15 -- JSON hyper-parameter loader with the reference's signature and return tuple."""
import json
import os
from typing import Tuple

DEFAULT_CONF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


def get_configs(env_name, agent_name, conf_dir=DEFAULT_CONF_DIR) -> Tuple[dict, dict, dict, dict, dict, dict]:
    try:
        file_location = os.path.join(conf_dir, env_name)
        with open(f"{file_location}.json") as json_file:
            params = json.load(json_file)
        return (params[agent_name]["agent_config"], params["trainer_config"], params["env_config"],
                params["tester_config"], params["tuner_config"], params)
    except KeyError:
        print("Can't find agent_config/trainer_config/env_config combo for the current problem! "
              "Check JSON file in configs dir!")
