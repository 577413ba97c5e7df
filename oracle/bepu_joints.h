// ORACLE — test infrastructure only (see bepu_math.h header). Joint / motor / servo / limit constraint functions.
#pragma once
#include "bepu_math.h"
#include "bepu_contacts.h"

namespace bepu_oracle {

template <class R> inline void register_joints(R& r) { (void)r; }

}  // namespace bepu_oracle
