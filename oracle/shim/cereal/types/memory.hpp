// stand-in for a cereal header: the reference's serialize()/save()/load() templates are never instantiated in oracle/_ref
#pragma once
